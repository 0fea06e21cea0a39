// CCollisionGeometry.h — the six axis-aligned walls of the simulation box.
//
// The reference digs the wall planes out of a Qt3D cuboid's vertex buffer
// (src/CCollisionGeometry.cpp:13-95) only to end up with +-extent/2 per axis
// (include/CCollisionGeometry.h:79-120).  Headless, the walls are built from the box size directly;
// the penalty response itself (inverseBoundingBoxBounce, src/CCollisionGeometry.cpp:117-133) runs on
// the device, fused into the integration kernel.
//
// Faces: the reference also extracts the triangles of the geometry (sVertex/sFace, include/CCollisionGeometry.h:
// 32-77) for inverseBounce (src/CCollisionGeometry.cpp:97-115), the per-face response "for a general object" that
// its step never calls.  Headless there is no Qt3D vertex buffer to read, so faces are set explicitly; the response
// runs on the device next to the wall term (sph_set_collision_faces).
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

struct sVertex {
    QVector3D m_pos, m_normal;
    sVertex() = default;
    explicit sVertex(QVector3D pos) : m_pos(pos) {}
    sVertex(QVector3D pos, QVector3D normal) : m_pos(pos), m_normal(normal) {}
};

struct sFace {
    sVertex m_v0, m_v1, m_v2;
    QVector3D m_normal;
    sFace() = default;
    // as in the reference: without an explicit normal the face takes the normal of its first vertex
    sFace(sVertex v0, sVertex v1, sVertex v2) : m_v0(v0), m_v1(v1), m_v2(v2), m_normal(v0.m_normal) {}
    sFace(sVertex v0, sVertex v1, sVertex v2, QVector3D normal) : m_v0(v0), m_v1(v1), m_v2(v2), m_normal(normal) {}
};

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
    const std::vector<sFace> &getFaces() const { return m_faces; }
    void setFaces(std::vector<sFace> faces) { m_faces = std::move(faces); }

private:
    sBoundingBox m_box;
    std::vector<sFace> m_faces;
};
