// sph_device.cuh — shared device-side definitions of the SPH step (sm_100a).
//
// Data layout in HBM (SoA, 16-byte vectors, everything resident between steps):
//   pos4[N]  = (x, y, z, id bits)      vel4[N] = (vx, vy, vz, 0)
//   dp4[N]   = (rho, p, p/rho^2, 1/rho) acc4[N] = (ax, ay, az, 0)
//   key[N]   = cell id, cell_start[cells+1] = exclusive scan of the per-cell counts
// "A" buffers hold last step's order, "S" buffers the canonical order (cell_id, id) of this step.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sph {

struct WallDev {
    float nx, ny, nz;  // normal
    float px, py, pz;  // position
};

// Kernel parameter block, passed by value (__grid_constant__).
struct Params {
    int rx, ry, rz;       // grid resolution (x fastest, z slowest: include/CGrid.h:27)
    int n_cells;          // rx*ry*rz
    int rz_global;        // z resolution of the whole tank (== rz on a single device)
    int z_base;           // slab mode: global index of local z-layer 0 (0 on a single device)
    double hbx, hby, hbz; // double(box)/2.0           (src/CCPUParticleSimulator.cpp:46-48)
    double h_d;           // double(h)
    float h, h2;          // h2 = fl32(h*h) == 0x3b08df0c
    float hbx_f, hby_f, hbz_f;  // float(half box): cell-face positions for the conservative row pruning
    float prune_margin;         // slack (1e-3 h) that makes the pruning bounds safe under fp32 rounding
    int tuning;                 // experiment switches (sph_set_option("tuning", bits)); 0 in production
    float neg_zero;             // -0.0f as a RUNTIME value: fma(d, d, -0) == fl(d*d), and ptxas cannot re-fuse it (see v2)
    float dt;
    float mass, viscosity, gas_stiffness, rest_density;
    float poly6_f, spiky_f, visc_f;  // fp64 coefficients rounded once to fp32
    float gx, gy, gz;                 // gravity
    float wall_k_f;                   // float(WALL_K)
    double wall_damping_d, wall_skin_d;
    int wall_count;
    WallDev walls[6];
};

// (dx*dx + dy*dy) + dz*dz with one rounding per operation — QVector3D::lengthSquared on a build
// without FMA contraction.  __fmul_rn/__fadd_rn are never fused by nvcc.
__device__ __forceinline__ float r2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// src/CCPUParticleSimulator.cpp:46-70: (int)floor((double(x) + double(b)/2.0)/double(h)), clamped.
__device__ __forceinline__ int cell_coord(float x, double half_box, double h_d, int res) {
    double q = __ddiv_rn(__dadd_rn((double)x, half_box), h_d);
    int c = __double2int_rd(q);
    return min(max(c, 0), res - 1);
}

__device__ __forceinline__ int cell_key(float4 p, const Params &P) {
    int cx = cell_coord(p.x, P.hbx, P.h_d, P.rx);
    int cy = cell_coord(p.y, P.hby, P.h_d, P.ry);
    // slab mode: the global layer (clamped like the reference clamps it) is shifted into the local grid
    int cz = cell_coord(p.z, P.hbz, P.h_d, P.rz_global) - P.z_base;
    cz = min(max(cz, 0), P.rz - 1);
    return cx + cy * P.rx + cz * P.rx * P.ry;
}

// Same enumeration, but the callback also gets the row slot (dz+1)*3 + (dy+1) in 0..8.
template <typename F>
__device__ __forceinline__ void for_each_row_slot(int key, const int *__restrict__ cell_start, const Params &P, F &&f) {
    const int rxy = P.rx * P.ry;
    const int cz = key / rxy;
    const int rem = key - cz * rxy;
    const int cy = rem / P.rx;
    const int cx = rem - cy * P.rx;
    const int xl = max(cx - 1, 0), xr = min(cx + 1, P.rx - 1);
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= P.rz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= P.ry) continue;
            const int c0 = xl + y * P.rx + z * rxy;
            const int a = __ldg(cell_start + c0);
            const int b = __ldg(cell_start + c0 + (xr - xl) + 1);
            f((dz + 1) * 3 + (dy + 1), a, b);
        }
    }
}

// Up to 9 contiguous candidate ranges (x-1..x+1 merged because x is the fastest cell axis).
// f(a, b) is called with the sorted-particle index range [a, b) of each (dy, dz) row, z-major.
template <typename F>
__device__ __forceinline__ void for_each_row(int key, const int *__restrict__ cell_start, const Params &P, F &&f) {
    const int rxy = P.rx * P.ry;
    const int cz = key / rxy;
    const int rem = key - cz * rxy;
    const int cy = rem / P.rx;
    const int cx = rem - cy * P.rx;
    const int xl = max(cx - 1, 0), xr = min(cx + 1, P.rx - 1);
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= P.rz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= P.ry) continue;
            const int c0 = xl + y * P.rx + z * rxy;
            const int a = __ldg(cell_start + c0);
            const int b = __ldg(cell_start + c0 + (xr - xl) + 1);
            f(a, b);
        }
    }
}

}  // namespace sph
