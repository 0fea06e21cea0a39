// sph_neighbours.cu — density/pressure and pressure+viscosity force kernels (sm_100a), plus the fused
// wall-collision + integration kernel.
//
// Counterparts of CCPUParticleSimulator::updateDensityPressure / updateForces / integrate
// (src/CCPUParticleSimulator.cpp:93-229) and CCollisionGeometry::inverseBoundingBoxBounce
// (src/CCollisionGeometry.cpp:117-133).  One thread per particle of the canonical (cell,id) order;
// the 27 neighbour cells collapse into <= 9 contiguous index ranges because x is the fastest cell
// axis, so candidate loads are 16-byte vector loads that are uniform or contiguous across a warp.
//
// The neighbour predicate is the un-contracted fp32 r2 <= h2 of the reference (bit-exact sets);
// everything after the predicate is plain fp32 (the reference mixes fp64 scalars in; rel 1e-5).
#include "sph_kernels.h"

namespace sph {

// ================================================================= density + pressure
// variant 0: candidates straight from global memory through L1 (LDG.128, read-only path).
__global__ void __launch_bounds__(128) k_density(const float4 *__restrict__ pos, const int *__restrict__ key,
                                                 const int *__restrict__ cell_start, float4 *__restrict__ dp,
                                                 int *__restrict__ nb_count, int n,
                                                 const __grid_constant__ Params P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 pi = __ldg(pos + i);
    const float h2 = P.h2;
    float sum = 0.0f;
    int cnt = 0;
    for_each_row(__ldg(key + i), cell_start, P, [&](int a, int b) {
#pragma unroll 4
        for (int j = a; j < b; ++j) {
            const float4 pj = __ldg(pos + j);
            const float r2 = r2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
            const float t = h2 - r2;  // sign(t) is exact: t >= 0  <=>  r2 <= h2
            if (t >= 0.0f) {
                ++cnt;
                sum = fmaf(t * t, t, sum);
            }
        }
    });
    // Wpoly6 summed, then rho *= mass; p = k (rho - rho0)   (src/CCPUParticleSimulator.cpp:9-15,133-134)
    float rho = sum * P.poly6_f;
    rho *= P.mass;
    const float prs = P.gas_stiffness * (rho - P.rest_density);
    const float inv_rho = 1.0f / rho;
    dp[i] = make_float4(rho, prs, prs * inv_rho * inv_rho, inv_rho);
    if (nb_count) nb_count[i] = cnt;
}

void launch_density(const float4 *pos_s, const int *key_s, const int *cell_start, float4 *dp, int *nb_count, int n,
                    const Params &P, cudaStream_t st) {
    if (n <= 0) return;
    k_density<<<(n + 127) / 128, 128, 0, st>>>(pos_s, key_s, cell_start, dp, nb_count, n, P);
}

// ================================================================= forces
// Per in-range pair (src/CCPUParticleSimulator.cpp:17-30,174-186):
//   gradW = spiky (h-r)^2 d / r ;  f_p += (p_i/rho_i^2 + p_j/rho_j^2) gradW
//   f_v  += (v_j - v_i) visc (h-r) / rho_j
struct ForceSum {
    float px, py, pz, vx, vy, vz;
};

__device__ __forceinline__ void force_pair(ForceSum &f, const float4 pi, const float4 vi, const float Ai, const float4 pj,
                                           const float4 vj, const float4 dj, const float r2, const Params &P) {
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const float inv_r = rsqrtf(r2);
    const float r = r2 * inv_r;
    const float hr = P.h - r;
    const float g = (P.spiky_f * hr) * (hr * inv_r) * (Ai + dj.z);
    f.px = fmaf(g, dx, f.px);
    f.py = fmaf(g, dy, f.py);
    f.pz = fmaf(g, dz, f.pz);
    const float l = (P.visc_f * hr) * dj.w;
    f.vx = fmaf(l, vj.x - vi.x, f.vx);
    f.vy = fmaf(l, vj.y - vi.y, f.vy);
    f.vz = fmaf(l, vj.z - vi.z, f.vz);
}

__device__ __forceinline__ float4 force_finish(const ForceSum &f, const float rho, const Params &P) {
    // f_p *= -m rho_i ; f_v *= mu m ; a = (f_p + f_v + g rho_i) / rho_i   (src/CCPUParticleSimulator.cpp:191-195)
    const float sp = -P.mass * rho, sv = P.viscosity * P.mass;
    float4 a;
    a.x = (f.px * sp + f.vx * sv + P.gx * rho) / rho;
    a.y = (f.py * sp + f.vy * sv + P.gy * rho) / rho;
    a.z = (f.pz * sp + f.vz * sv + P.gz * rho) / rho;
    a.w = 0.0f;
    return a;
}

// variant 0: scan candidates cheaply, queue the hits per thread in shared memory, then run the
// expensive pair body densely over the queue (avoids paying the body for every candidate the
// moment any lane of the warp has a hit).
template <int BLOCK, int QCAP>
__global__ void __launch_bounds__(BLOCK) k_forces(const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                                  const float4 *__restrict__ dp, const int *__restrict__ key,
                                                  const int *__restrict__ cell_start, float4 *__restrict__ acc, int n,
                                                  const __grid_constant__ Params P) {
    __shared__ int queue[QCAP][BLOCK];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * BLOCK + tid;
    if (i >= n) return;
    const float4 pi = __ldg(pos + i);
    const float4 vi = __ldg(vel + i);
    const float4 di = __ldg(dp + i);
    const float h2 = P.h2;
    ForceSum f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int qn = 0;
    for_each_row(__ldg(key + i), cell_start, P, [&](int a, int b) {
#pragma unroll 4
        for (int j = a; j < b; ++j) {
            const float4 pj = __ldg(pos + j);
            const float r2 = r2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
            if (h2 - r2 >= 0.0f && j != i) {
                if (qn < QCAP) {
                    queue[qn++][tid] = j;
                } else {  // queue full: process in place (rare, very dense neighbourhoods)
                    force_pair(f, pi, vi, di.z, pj, __ldg(vel + j), __ldg(dp + j), r2, P);
                }
            }
        }
    });
    for (int e = 0; e < qn; ++e) {
        const int j = queue[e][tid];
        const float4 pj = __ldg(pos + j);
        const float r2 = r2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        force_pair(f, pi, vi, di.z, pj, __ldg(vel + j), __ldg(dp + j), r2, P);
    }
    acc[i] = force_finish(f, di.x, P);
}

void launch_forces(const float4 *pos_s, const float4 *vel_s, const float4 *dp, const int *key_s, const int *cell_start,
                   float4 *acc, int n, const Params &P, cudaStream_t st) {
    if (n <= 0) return;
    k_forces<128, 48><<<(n + 127) / 128, 128, 0, st>>>(pos_s, vel_s, dp, key_s, cell_start, acc, n, P);
}

// ================================================================= walls + integration (fused)
// Wall penalty force exactly as the CPU path computes it (fp32 vectors, fp64 scalar d), accumulator
// starting at 0 (the OpenCL kernel's (1,1,0) start, resources/kernels/sph_common.cl:68, is a bug the
// CPU path does not have).  Then x' = (x + v dt) + (a dt) dt ; v' = (x' - x)/dt with one rounding per
// operation (src/CCPUParticleSimulator.cpp:220-221) — bit-exact given the same acceleration.
__global__ void __launch_bounds__(256) k_integrate_collide(const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                                           float4 *__restrict__ acc, float4 *__restrict__ pos_out,
                                                           float4 *__restrict__ vel_out, int i0, int n,
                                                           const int *__restrict__ key, int *__restrict__ far_movers,
                                                           const __grid_constant__ Params P) {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;  // particles [i0, n): slab mode runs sub-ranges
    if (i >= n) return;
    const float4 p = __ldg(pos + i);
    const float4 v = __ldg(vel + i);
    float4 a = acc[i], np, nv;
    walls_and_integrate(p, v, a, np, nv, P);
    if (far_movers) {
        // slab mode: count the particles the boundary-only exchange would miss (see exchange_would_miss)
        const int old_layer = __ldg(key + i) / (P.rx * P.xb * P.ry) + P.z_base;
        const int new_layer = cell_coord(np.z, P.hbz, P.h_d, P.rz_global);
        if (exchange_would_miss(old_layer, new_layer, P)) atomicAdd(far_movers, 1);
    }
    pos_out[i] = np;  // written to the authoritative A buffers, in the canonical order of this step
    vel_out[i] = nv;
    acc[i] = a;  // total acceleration (SPH + walls), what updateForces leaves in the CPU path
}

void launch_integrate_collide(const float4 *pos_s, const float4 *vel_s, float4 *acc, float4 *pos_out, float4 *vel_out,
                              int i0, int i1, const int *key_s, int *far_movers, const Params &P, cudaStream_t st) {
    if (i1 <= i0) return;
    k_integrate_collide<<<(i1 - i0 + 255) / 256, 256, 0, st>>>(pos_s, vel_s, acc, pos_out, vel_out, i0, i1, key_s,
                                                               far_movers, P);
}

}  // namespace sph
