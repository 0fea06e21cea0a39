// CGrid.h — uniform grid descriptor (resolution, cell numbering, walls).
//
// In the reference CGrid owns one std::vector<CParticle*> per cell (include/CGrid.h:24-28,
// src/CGrid.cpp:17-18) and is a Qt3D wireframe entity.  Here the per-cell storage lives on the device
// (cell_start[] over particles sorted by cell id), so the host object only carries the geometry: the
// cell id is x + y*ResX + z*ResX*ResY exactly as CGrid::at() computes it.
#pragma once

#include "CCollisionGeometry.h"
#include "QtCompat.h"

class CGrid {
public:
    CGrid(const QVector3D &size, const QVector3D &resolution)
        : m_ResX((int)resolution.x()), m_ResY((int)resolution.y()), m_ResZ((int)resolution.z()),
          m_cell_count(m_ResX * m_ResY * m_ResZ), m_collisionGeometry(size) {}

    int xRes() const { return m_ResX; }
    int yRes() const { return m_ResY; }
    int zRes() const { return m_ResZ; }
    const int &getCellCount() const { return m_cell_count; }
    int cellId(int x, int y, int z) const { return x + y * m_ResX + z * m_ResX * m_ResY; }
    CCollisionGeometry *getCollisionGeometry() { return &m_collisionGeometry; }

private:
    int m_ResX, m_ResY, m_ResZ, m_cell_count;
    CCollisionGeometry m_collisionGeometry;
};
