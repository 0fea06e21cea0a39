// CCollisionGeometry.h — the six axis-aligned walls of the simulation box.
//
// The reference digs the wall planes out of a Qt3D cuboid's vertex buffer
// (src/CCollisionGeometry.cpp:13-95) only to end up with +-extent/2 per axis
// (include/CCollisionGeometry.h:79-120).  Headless, the walls are built from the box size directly;
// the penalty response itself (inverseBoundingBoxBounce, src/CCollisionGeometry.cpp:117-133) runs on
// the device, fused into the integration kernel.
#pragma once

#include <vector>

#include "CParticle.h"
#include "QtCompat.h"

#define WALL_K 10000.0
#define WALL_DAMPING (-0.9)

struct sWall {
    cl_float3 normal;
    cl_float3 position;
};
static_assert(sizeof(sWall) == 32, "sWall must stay 32 bytes");

struct sBoundingBox {
    QVector3D m_min, m_max;
    std::vector<sWall> m_walls;  // left, bottom, back, right, top, front
};

class CCollisionGeometry {
public:
    explicit CCollisionGeometry(const QVector3D &extent) {
        m_box.m_max = QVector3D(extent.x() / 2.0f, extent.y() / 2.0f, extent.z() / 2.0f);
        m_box.m_min = -m_box.m_max;
        const float lo[3] = {m_box.m_min.x(), m_box.m_min.y(), m_box.m_min.z()};
        const float hi[3] = {m_box.m_max.x(), m_box.m_max.y(), m_box.m_max.z()};
        m_box.m_walls.resize(6);
        for (int axis = 0; axis < 3; ++axis) {
            float nlo[3] = {0, 0, 0}, plo[3] = {0, 0, 0}, nhi[3] = {0, 0, 0}, phi[3] = {0, 0, 0};
            nlo[axis] = -1.f; plo[axis] = lo[axis];
            nhi[axis] = 1.f;  phi[axis] = hi[axis];
            m_box.m_walls[axis] = {{nlo[0], nlo[1], nlo[2], 0.f}, {plo[0], plo[1], plo[2], 0.f}};
            m_box.m_walls[axis + 3] = {{nhi[0], nhi[1], nhi[2], 0.f}, {phi[0], phi[1], phi[2], 0.f}};
        }
    }
    const sBoundingBox &getBoundingBox() const { return m_box; }

private:
    sBoundingBox m_box;
};
