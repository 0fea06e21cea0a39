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
    int xb;               // x bins per cell of the sort key (power of two; 1 = sort by the reference cell id only)
    double h_win;         // h (1 + 1e-6): half-width of the candidate x-window
    int rz_global;        // z resolution of the whole tank (== rz on a single device)
    int z_base;           // slab mode: global index of local z-layer 0 (0 on a single device)
    int own_z0, own_z1;   // slab mode: owned global layers [own_z0, own_z1) (0 and rz_global on a single device)
    int next_z0, next_z1; // slab mode: the owned layers AFTER this step's exchange (the faces may move by load, see Slab)
    int kz_lo, kz_hi;     // local z-layers [kz_lo, kz_hi) a sort key may name (0 and rz on a single device; in slab mode the
                          // owned + ghost layers and one more on either side: only those bins are scanned, see launch_scan)
    int dens_key_lo, dens_key_hi;  // sort keys [lo, hi) whose density is needed (everything on a single device; in slab
                                   // mode the owned layers and ONE ghost layer per face: the outer ghost layer only lends
                                   // its positions to the inner one)
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
    // optional collision mesh (sph_set_collision_faces): plane k = {planes[2k] = inverse unit normal, planes[2k+1] =
    // vertex}, three planes per face in face/vertex order; 0 planes = the shipped behaviour
    const float4 *mesh_planes;
    int mesh_plane_count;
    float mesh_k_f;  // float(5000.0)
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

// ---- sort keys ---------------------------------------------------------------------------------------
// The device sorts by a FINE key: every reference cell is split into P.xb (1, 2, 4 or 8) bins along x,
//   fine key = x_bin + cy * (rx*xb) + cz * (rx*xb) * ry,      x_bin = clamp(floor(q_x * xb), 0, rx*xb - 1),
// with q_x = (double(x) + b/2)/h exactly as the reference computes it.  q_x * xb is exact for a power of two, so
// x_bin / xb == clamp(floor(q_x)) is the reference's cell coordinate bit for bit and the reference cell id
// (include/CGrid.h:27) is recovered with coarse_key().  The order (cell, x_bin, id) refines the cell order: a
// reference cell is still one contiguous particle range, and inside a row the particles are roughly sorted by x, so
// the density pass can scan an x-WINDOW of about (2 xb + 1)/(3 xb) of the row instead of all three cells.
__device__ __forceinline__ int x_bin_unclamped(double x, const Params &P) {
    return __double2int_rd(__dmul_rn(__ddiv_rn(__dadd_rn(x, P.hbx), P.h_d), (double)P.xb));
}

__device__ __forceinline__ int cell_key(float4 p, const Params &P) {
    const int rxb = P.rx * P.xb;
    const int gx = min(max(x_bin_unclamped((double)p.x, P), 0), rxb - 1);
    const int cy = cell_coord(p.y, P.hby, P.h_d, P.ry);
    // slab mode: the global layer (clamped like the reference clamps it) is shifted into the local grid
    int cz = cell_coord(p.z, P.hbz, P.h_d, P.rz_global) - P.z_base;
    cz = min(max(cz, P.kz_lo), P.kz_hi - 1);
    return gx + cy * rxb + cz * rxb * P.ry;
}

struct CellCoords {
    int gx, cx, cy, cz;  // x bin, reference cell coordinates (cz local in slab mode)
};
__device__ __forceinline__ CellCoords decode_key(int key, const Params &P) {
    const int rxb = P.rx * P.xb, plane = rxb * P.ry;
    CellCoords c;
    c.cz = key / plane;
    const int rem = key - c.cz * plane;
    c.cy = rem / rxb;
    c.gx = rem - c.cy * rxb;
    c.cx = c.gx / P.xb;
    return c;
}
// reference cell id x + y*ResX + z*ResX*ResY (z local in slab mode)
__device__ __forceinline__ int coarse_key(int key, const Params &P) {
    const CellCoords c = decode_key(key, P);
    return c.cx + c.cy * P.rx + c.cz * P.rx * P.ry;
}

// Up to 9 contiguous candidate ranges: the reference's 27 cells, x-1..x+1 merged because x is the fastest axis.
// f(slot, a, b) gets the row slot (dz+1)*3 + (dy+1) and the sorted-particle index range [a, b), z-major.
template <typename F>
__device__ __forceinline__ void for_each_row_slot(int key, const int *__restrict__ cell_start, const Params &P, F &&f) {
    const CellCoords c = decode_key(key, P);
    const int rxb = P.rx * P.xb, plane = rxb * P.ry;
    const int xl = max(c.cx - 1, 0), xr = min(c.cx + 1, P.rx - 1);
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = c.cz + dz;
        if (z < 0 || z >= P.rz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = c.cy + dy;
            if (y < 0 || y >= P.ry) continue;
            const int row = y * rxb + z * plane;
            const int a = __ldg(cell_start + row + xl * P.xb);
            const int b = __ldg(cell_start + row + (xr + 1) * P.xb);
            f((dz + 1) * 3 + (dy + 1), a, b);
        }
    }
}

template <typename F>
__device__ __forceinline__ void for_each_row(int key, const int *__restrict__ cell_start, const Params &P, F &&f) {
    for_each_row_slot(key, cell_start, P, [&](int, int a, int b) { f(a, b); });
}

// The same rows clipped to the x-window a particle at x = px can have neighbours in: bins
// [bin(px - h'), bin(px + h')] with h' = h (1 + 1e-6), intersected with the reference's three cells.  A candidate j
// that passes the fp32 predicate has |x_i - x_j| <= h (1 + 2^-22) in exact arithmetic and q_x is evaluated by the
// same monotone fp64 expression for both, so j's bin always lies inside the window: the neighbour sets do not
// depend on xb.
template <typename F>
__device__ __forceinline__ void for_each_window_slot(float px, int key, const int *__restrict__ cell_start,
                                                     const Params &P, F &&f) {
    const CellCoords c = decode_key(key, P);
    const int rxb = P.rx * P.xb, plane = rxb * P.ry;
    const int xl = max(c.cx - 1, 0), xr = min(c.cx + 1, P.rx - 1);
    const int bl = max(x_bin_unclamped(__dsub_rn((double)px, P.h_win), P), xl * P.xb);
    const int bh = min(x_bin_unclamped(__dadd_rn((double)px, P.h_win), P), (xr + 1) * P.xb - 1);
    // a clamped (out-of-box) particle sits in an edge bin whatever its coordinate says: keep its own bin inside
    const int lo = min(bl, c.gx), hi = max(bh, c.gx);
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = c.cz + dz;
        if (z < 0 || z >= P.rz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = c.cy + dy;
            if (y < 0 || y >= P.ry) continue;
            const int row = y * rxb + z * plane;
            const int a = __ldg(cell_start + row + lo);
            const int b = __ldg(cell_start + row + hi + 1);
            f((dz + 1) * 3 + (dy + 1), a, b);
        }
    }
}

// Slab mode: the overlapped exchange packs only the owned layers next to each interior face: from the face this
// rank owns now down to four layers inside the face it will own after the exchange (two that become the neighbour's
// ghosts plus two of slack; the face may move, see Slab).  A particle that starts further inside and ends the step
// in the outer two layers or beyond would be missed as a ghost / migrant; such particles are counted
// ("slab_far_movers") and the count must stay 0 (it takes more than two layers = 0.09 m per step, i.e. more than
// 9 m/s towards the face).
__device__ __forceinline__ bool exchange_would_miss(int old_layer, int new_layer, const Params &P) {
    return (P.own_z0 > 0 && old_layer >= max(P.own_z0, P.next_z0) + 4 && new_layer < P.next_z0 + 2) ||
           (P.own_z1 < P.rz_global && old_layer < min(P.own_z1, P.next_z1) - 4 && new_layer >= P.next_z1 - 2);
}

// ---- walls + integration, shared by k_integrate_collide and the fused force kernel -------------------
__device__ __forceinline__ float dot_exact(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// Wall penalty force exactly as the CPU path computes it (fp32 vectors, fp64 scalar d, accumulator from 0;
// src/CCollisionGeometry.cpp:117-133) added to the SPH acceleration `a`, then x' = (x + v dt) + (a dt) dt and
// v' = (x' - x)/dt with one rounding per operation (src/CCPUParticleSimulator.cpp:220-221).
__device__ __forceinline__ void walls_and_integrate(const float4 p, const float4 v, float4 &a, float4 &np, float4 &nv,
                                                    const Params &P) {
    float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
    for (int w = 0; w < 6; ++w) {
        if (w >= P.wall_count) break;
        const WallDev W = P.walls[w];
        const float inx = __fmul_rn(W.nx, -1.0f), iny = __fmul_rn(W.ny, -1.0f), inz = __fmul_rn(W.nz, -1.0f);
        const double d = __dadd_rn((double)dot_exact(__fsub_rn(W.px, p.x), __fsub_rn(W.py, p.y), __fsub_rn(W.pz, p.z), inx, iny, inz),
                                   P.wall_skin_d);
        if (d > 0.0) {
            const float df = (float)d;
            wx = __fadd_rn(wx, __fmul_rn(__fmul_rn(P.wall_k_f, inx), df));
            wy = __fadd_rn(wy, __fmul_rn(__fmul_rn(P.wall_k_f, iny), df));
            wz = __fadd_rn(wz, __fmul_rn(__fmul_rn(P.wall_k_f, inz), df));
            const float s = (float)__dmul_rn(P.wall_damping_d, (double)dot_exact(v.x, v.y, v.z, inx, iny, inz));
            wx = __fadd_rn(wx, __fmul_rn(s, inx));
            wy = __fadd_rn(wy, __fmul_rn(s, iny));
            wz = __fadd_rn(wz, __fmul_rn(s, inz));
        }
    }
    a.x = __fadd_rn(a.x, wx);
    a.y = __fadd_rn(a.y, wy);
    a.z = __fadd_rn(a.z, wz);
    if (P.mesh_plane_count > 0) {
        // CCollisionGeometry::inverseBounce (src/CCollisionGeometry.cpp:97-115): same shape as the wall term, one
        // plane per face vertex, spring 5000.0 and the literals -0.9 / 0.01; its own accumulator starts at 0
        float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll 1
        for (int k = 0; k < P.mesh_plane_count; ++k) {
            const float4 in = __ldg(P.mesh_planes + 2 * k), q = __ldg(P.mesh_planes + 2 * k + 1);
            const double d = __dadd_rn((double)dot_exact(__fsub_rn(q.x, p.x), __fsub_rn(q.y, p.y), __fsub_rn(q.z, p.z), in.x, in.y, in.z), 0.01);
            if (d > 0.0) {
                const float df = (float)d;
                mx = __fadd_rn(mx, __fmul_rn(__fmul_rn(P.mesh_k_f, in.x), df));
                my = __fadd_rn(my, __fmul_rn(__fmul_rn(P.mesh_k_f, in.y), df));
                mz = __fadd_rn(mz, __fmul_rn(__fmul_rn(P.mesh_k_f, in.z), df));
                const float s = (float)__dmul_rn(-0.9, (double)dot_exact(v.x, v.y, v.z, in.x, in.y, in.z));
                mx = __fadd_rn(mx, __fmul_rn(s, in.x));
                my = __fadd_rn(my, __fmul_rn(s, in.y));
                mz = __fadd_rn(mz, __fmul_rn(s, in.z));
            }
        }
        a.x = __fadd_rn(a.x, mx);
        a.y = __fadd_rn(a.y, my);
        a.z = __fadd_rn(a.z, mz);
    }
    const float dt = P.dt;
    np.x = __fadd_rn(__fadd_rn(p.x, __fmul_rn(v.x, dt)), __fmul_rn(__fmul_rn(a.x, dt), dt));
    np.y = __fadd_rn(__fadd_rn(p.y, __fmul_rn(v.y, dt)), __fmul_rn(__fmul_rn(a.y, dt), dt));
    np.z = __fadd_rn(__fadd_rn(p.z, __fmul_rn(v.z, dt)), __fmul_rn(__fmul_rn(a.z, dt), dt));
    np.w = p.w;
    nv.x = __fdiv_rn(__fsub_rn(np.x, p.x), dt);
    nv.y = __fdiv_rn(__fsub_rn(np.y, p.y), dt);
    nv.z = __fdiv_rn(__fsub_rn(np.z, p.z), dt);
    nv.w = 0.0f;
}

}  // namespace sph
