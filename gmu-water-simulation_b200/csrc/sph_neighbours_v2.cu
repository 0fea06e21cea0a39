// sph_neighbours_v2.cu — the production neighbour passes (sm_100a):
//
//   k_density_mask : one scan of the 27 candidate cells (<= 9 contiguous rows) per particle.  Positions
//                    come from split SoA arrays (xs/ys/zs) with 16-byte loads = 4 candidates per load;
//                    the exact, un-contracted fp32 predicate r2 <= h2 is evaluated for two candidates
//                    per instruction with Blackwell's packed fp32x2 ops (FADD2/FFMA2, each element
//                    rounded to nearest: bit-identical to the scalar sequence).  Hits are recorded as
//                    a BITMASK: the sign bit of t = h2 - r2 is funnel-shifted into a 32-candidate word
//                    (one SHF per candidate, no predicate -> bit conversion, no list append), and the
//                    poly6 density is accumulated branch-free.  Each finished word is stored with the
//                    index of its first candidate.  Also writes the packed record
//                    fdat[i] = {x,y,z,p/rho^2 | vx,vy,vz,1/rho} the force pass gathers.
//   k_forces_mask  : walks the set bits of the particle's words as ONE flat loop (no per-row or
//                    per-word lock-step across the warp) and evaluates the pressure + viscosity pair
//                    term for true neighbours only; each neighbour is ONE 256-bit gather
//                    (LDG.E.ENL2.256, sm_100+).
//
// Rows start at an index aligned down to 4 and end aligned up; the bits of the <= 3 foreign slots on either
// side are dropped from the word, and a word in which a foreign slot was within h is rolled back and re-scanned
// with those slots masked, so counts, sets and the density sum are exact and independent of the array offset.  (Per-particle geometric pruning of rows/cells was measured and
// dropped: under SIMT a warp pays for its longest lane, so it only lowered lane utilisation.  Staging the
// candidate rows in shared memory with bulk async copies was built and measured in round 1 — slower, the rows
// are served well enough by L1 — and removed; DESIGN.md §3.2 keeps the numbers.)
#include "sph_kernels.h"

namespace sph {

typedef unsigned long long u64;

// ---- packed fp32x2 helpers (PTX ISA 8.6+, sm_100+) --------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// h2 - ((dx*dx + dy*dy) + dz*dz) for two candidates, every operation rounded separately.
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry an explicit .rn
// (and regardless of -fmad=false), which would change the predicate.  The squares are therefore formed
// as fma(d, d, -0.0) with the -0.0 coming from a kernel parameter: one rounding of the exact product,
// i.e. fl(d*d) bit for bit, and an FFMA2 result cannot be contracted into the following FADD2.
__device__ __forceinline__ u64 t_exact2(u64 px, u64 py, u64 pz, u64 x, u64 y, u64 z, u64 h2, u64 nz) {
    const u64 dx = sub2(px, x), dy = sub2(py, y), dz = sub2(pz, z);
    return sub2(h2, add2(add2(fma2(dx, dx, nz), fma2(dy, dy, nz)), fma2(dz, dz, nz)));
}

// MUFU.RSQ without the denormal-rescaling wrapper rsqrtf() carries (r2 is never that small for a real pair)
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void ld256(const float4 *p, float4 &a, float4 &b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

// ================================================================= density + pressure + hit bitmask
// Word format: bit 31 = first candidate of the word (index j0, a multiple of 4), bit 31-k = candidate j0+k.
// Stored as uint2 {bits, j0 + 31} so that the force pass gets j = (j0 + 31) - msb_index(bits).
//
// Every row [a, b) of the particle's x-window is scanned in aligned groups of 4 candidates, 8 groups = one word.
// The first and the last group of a row may contain FOREIGN slots (the sorted array continues with other bins
// there).  All lanes of a warp run the same unmasked loop; the foreign bits are dropped when the word is finished.
// A foreign slot that happens to lie within h (possible next to empty cells, rare) would also have added its
// poly6 term to the sum: such a word is detected by its dropped bits, the accumulator is rolled back to its value at
// the start of the word and the word is re-scanned with the foreign slots moved to a far-away sentinel — so the
// density sum only ever contains the candidates of [a, b), each added exactly once, in ascending order.
//
// The sum lives in ONE packed accumulator {even, odd} indexed by the parity of the candidate's position INSIDE ITS
// ROW (j - a), not by its absolute index: the halves are swapped whenever the parity of the row start changes.
// Together with the roll-back this makes the density a pure function of the particle state — it does not depend on
// where the sorted array happens to start, so a slab of a multi-GPU run (whose local array has a different offset)
// produces the same bits as the single-GPU run (tests/test_slab.py: k slabs == 1 GPU bit for bit).
constexpr float kFar = 1.0e18f;  // sentinel coordinate: (kFar)^2 is finite in fp32 and t = h2 - r2 hugely negative

struct DensityScan {
    u64 px2, py2, pz2, h22, nz2;
    u64 acc;        // packed poly6 sum {even-in-row, odd-in-row}
    unsigned miss;  // sign bits of t for the current word, first candidate in the highest bit shifted in
};

// MASKED: slots k of the group outside [lo, hi) are foreign and moved out of reach before the arithmetic
template <bool MASKED>
__device__ __forceinline__ void scan_group(DensityScan &d, const float *__restrict__ xs, const float *__restrict__ ys,
                                           const float *__restrict__ zs, const int j, const int lo, const int hi) {
    ulonglong2 X = __ldg(reinterpret_cast<const ulonglong2 *>(xs + j));
    const ulonglong2 Y = __ldg(reinterpret_cast<const ulonglong2 *>(ys + j));
    const ulonglong2 Z = __ldg(reinterpret_cast<const ulonglong2 *>(zs + j));
    if (MASKED) {
        float x0, x1, x2, x3;
        upk(X.x, x0, x1);
        upk(X.y, x2, x3);
        x0 = (0 >= lo && 0 < hi) ? x0 : kFar;
        x1 = (1 >= lo && 1 < hi) ? x1 : kFar;
        x2 = (2 >= lo && 2 < hi) ? x2 : kFar;
        x3 = (3 >= lo && 3 < hi) ? x3 : kFar;
        X.x = pk(x0, x1);
        X.y = pk(x2, x3);
    }
    const u64 t01 = t_exact2(d.px2, d.py2, d.pz2, X.x, Y.x, Z.x, d.h22, d.nz2);
    const u64 t23 = t_exact2(d.px2, d.py2, d.pz2, X.y, Y.y, Z.y, d.h22, d.nz2);
    float t0, t1, t2, t3;
    upk(t01, t0, t1);
    upk(t23, t2, t3);
    // t >= 0  <=>  r2 <= h2, and sign(t) is exact (t is never -0): append the sign bits
    d.miss = __funnelshift_l(__float_as_uint(t0), d.miss, 1);
    d.miss = __funnelshift_l(__float_as_uint(t1), d.miss, 1);
    d.miss = __funnelshift_l(__float_as_uint(t2), d.miss, 1);
    d.miss = __funnelshift_l(__float_as_uint(t3), d.miss, 1);
    // poly6: sum += max(t,0)^3, branch-free; candidates in ascending order, even / odd slots
    const u64 c01 = pk(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), c23 = pk(fmaxf(t2, 0.0f), fmaxf(t3, 0.0f));
    d.acc = fma2(c01, mul2(c01, c01), d.acc);
    d.acc = fma2(c23, mul2(c23, c23), d.acc);
}

__global__ void __launch_bounds__(128, 8) k_density_mask(const float *__restrict__ xs, const float *__restrict__ ys,
                                                      const float *__restrict__ zs, const float4 *__restrict__ vel,
                                                      const int *__restrict__ key, const int *__restrict__ cell_start,
                                                      float4 *__restrict__ dp, float4 *__restrict__ fdat,
                                                      uint2 *__restrict__ mask, int *__restrict__ nb_count,
                                                      int *__restrict__ nb_words, int *__restrict__ ovf, int n,
                                                      const __grid_constant__ Params P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = __ldg(xs + i), py = __ldg(ys + i), pz = __ldg(zs + i);
    DensityScan d;
    d.px2 = pk(px, px);
    d.py2 = pk(py, py);
    d.pz2 = pk(pz, pz);
    d.h22 = pk(P.h2, P.h2);
    d.nz2 = pk(P.neg_zero, P.neg_zero);
    d.acc = 0ull;  // (+0.0f, +0.0f)
    d.miss = 0u;
    uint2 *const wbase = mask + ((size_t)(i >> 5) * kMaskWords) * 32 + (i & 31);
    int cnt = 0, widx = 0;
    int parity = 0;  // parity of the row start the accumulator halves are currently aligned with
    const int my_key = __ldg(key + i);
    // slab mode: the outer ghost layer of a face only lends its positions to the inner one; nobody reads its density
    const bool wanted = my_key >= P.dens_key_lo && my_key < P.dens_key_hi;
    if (wanted) for_each_window_slot(px, my_key, cell_start, P, [&](const int, const int a, const int b) {
        if (a >= b) return;
        if ((a ^ parity) & 1) {  // {even, odd} is relative to the row start: swap the halves when its parity flips
            float e, o;
            upk(d.acc, e, o);
            d.acc = pk(o, e);
            parity = a;
        }
        int j = a & ~3;
        while (j < b) {
            const int jw = j;
            const int jend = min(jw + 32, b);
            const u64 acc_at_word_start = d.acc;
            d.miss = 0u;
#pragma unroll 2
            for (; j < jend; j += 4) scan_group<false>(d, xs, ys, zs, j, 0, 4);
            // left-justify (a short last word shifted in fewer than 32 bits), turn misses into hits and keep
            // only the candidates inside [a, b)
            const int scanned = j - jw;  // multiple of 4, 4..32
            unsigned m = ~(d.miss << (32 - scanned));
            const int lo = max(a - jw, 0), hi = min(b - jw, 32);  // candidate offsets of this word inside the row
            const unsigned vm = (0xffffffffu >> lo) & ~((hi < 32) ? (0xffffffffu >> hi) : 0u);
            const unsigned foreign_hits = m & ~vm & ~((scanned < 32) ? (0xffffffffu >> scanned) : 0u);
            m &= vm;
            if (foreign_hits) {
                // rare: a foreign slot within h has been added to the sum.  Roll the word back and scan it again with
                // the foreign slots out of reach (same candidates, same order, now only those of [a, b)).
                d.acc = acc_at_word_start;
                for (int jj = jw; jj < jw + scanned; jj += 4) scan_group<true>(d, xs, ys, zs, jj, a - jj, b - jj);
            }
            cnt += __popc(m);
            if (m) {  // empty words are not stored at all
                if (widx < kMaskWords) wbase[widx * 32] = make_uint2(m, (unsigned)(jw + 31));
                ++widx;
            }
        }
    });
    float s_even, s_odd;
    upk(d.acc, s_even, s_odd);
    float sum = s_even + s_odd;  // commutative: independent of which half is which
    if (!(px < 1.0e17f) || !wanted) {
        // a particle with a non-finite coordinate (sentinel in xs/ys/zs): in the reference every comparison with NaN
        // is false, so it has no neighbours at all, not even itself -> density 0 (src/CCPUParticleSimulator.cpp:122-127).
        // A skipped outer-ghost particle gets the same harmless record.
        sum = 0.0f;
        cnt = 0;
        widx = 0;
    }
    // Wpoly6 summed, then rho *= mass; p = k (rho - rho0)   (src/CCPUParticleSimulator.cpp:9-15,133-134)
    float rho = sum * P.poly6_f;
    rho *= P.mass;
    const float prs = P.gas_stiffness * (rho - P.rest_density);
    const float inv_rho = 1.0f / rho;
    const float A = prs * inv_rho * inv_rho;
    __stcs(dp + i, make_float4(rho, prs, A, inv_rho));
    const float4 v = __ldg(vel + i);
    __stcs(fdat + 2 * (size_t)i, make_float4(px, py, pz, A));
    __stcs(fdat + 2 * (size_t)i + 1, make_float4(v.x, v.y, v.z, inv_rho));
    __stcs(nb_count + i, cnt);
    __stcs(nb_words + i, widx);  // > kMaskWords: the force pass takes the overflow path for this particle
    if (widx > kMaskWords) ovf[1 + atomicAdd(ovf, 1)] = i;
}

void launch_density_mask(const NbBuffers &nb, const float4 *vel_s, const int *key_s, const int *cell_start, float4 *dp,
                         int *nb_count, int n, const Params &P, cudaStream_t st, bool reset_overflow_list) {
    if (n <= 0) return;
    // k_rank_scatter leaves the list empty; only a second density pass on the same grid needs the explicit reset
    if (reset_overflow_list) cudaMemsetAsync(nb.ovf, 0, sizeof(int), st);
    k_density_mask<<<(n + 127) / 128, 128, 0, st>>>(nb.xs, nb.ys, nb.zs, vel_s, key_s, cell_start, dp, nb.fdat, nb.mask,
                                                    nb_count, nb.words, nb.ovf, n, P);
}

// ================================================================= forces from the bitmask
struct ForceSum {
    float px, py, pz, vx, vy, vz;
};

// Pair term (src/CCPUParticleSimulator.cpp:17-30,174-186).  `self` zeroes 1/r so the particle's own
// bit (r = 0) contributes nothing.
__device__ __forceinline__ void pair_term(ForceSum &f, const float4 pi, const float4 vi, const float4 pj, const float4 vj,
                                          const bool self, const Params &P) {
    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float inv_r = self ? 0.0f : rsqrt_fast(r2);
    const float r = r2 * inv_r;
    const float hr = P.h - r;
    const float g = (P.spiky_f * hr) * (hr * inv_r) * (pi.w + pj.w);
    f.px = fmaf(g, dx, f.px);
    f.py = fmaf(g, dy, f.py);
    f.pz = fmaf(g, dz, f.pz);
    const float l = (P.visc_f * hr) * vj.w;
    f.vx = fmaf(l, vj.x - vi.x, f.vx);
    f.vy = fmaf(l, vj.y - vi.y, f.vy);
    f.vz = fmaf(l, vj.z - vi.z, f.vz);
}

// f_p *= -m rho_i ; f_v *= mu m ; a = (f_p + f_v + g rho_i) / rho_i   (src/CCPUParticleSimulator.cpp:191-195)
__device__ __forceinline__ float4 force_result(const ForceSum &f, const float rho, const Params &P) {
    const float sp = -P.mass * rho, sv = P.viscosity * P.mass;
    return make_float4((f.px * sp + f.vx * sv + P.gx * rho) / rho, (f.py * sp + f.vy * sv + P.gy * rho) / rho,
                       (f.pz * sp + f.vz * sv + P.gz * rho) / rho, 0.0f);
}

// Epilogue of both force kernels.  FUSED (whole-step path): the wall term and the integration run right here -
// the particle's position and velocity are already in registers (its own fdat record) - and the new state goes
// to the A buffers; the separate integrate kernel and a round trip of acc through HBM disappear.  Bitwise the
// same result as forces + k_integrate_collide.
template <bool FUSED>
__device__ __forceinline__ void force_epilogue(const int i, const ForceSum &f, const float4 pi, const float4 vi,
                                               const float4 *__restrict__ dp, const float4 *__restrict__ pos_s,
                                               float4 *__restrict__ acc, float4 *__restrict__ pos_out,
                                               float4 *__restrict__ vel_out, const int *__restrict__ key,
                                               int *__restrict__ far_movers, const Params &P) {
    float4 a = force_result(f, __ldg(&dp[i].x), P);
    if (FUSED) {
        // the particle's real position: pos_s, not its candidate record (a non-finite particle carries the far-away
        // sentinel there and must stay non-finite, like in the phase path and in the reference); .w = particle id
        const float4 p = __ldg(pos_s + i);
        const float4 v = make_float4(vi.x, vi.y, vi.z, 0.0f);
        float4 np, nv;
        walls_and_integrate(p, v, a, np, nv, P);
        if (far_movers) {
            // slab mode: count the particles the boundary-only exchange would miss (see exchange_would_miss)
            const int old_layer = __ldg(key + i) / (P.rx * P.xb * P.ry) + P.z_base;
            const int new_layer = cell_coord(np.z, P.hbz, P.h_d, P.rz_global);
            if (exchange_would_miss(old_layer, new_layer, P)) atomicAdd(far_movers, 1);
        }
        pos_out[i] = np;
        vel_out[i] = nv;
    }
    acc[i] = a;
}

// Particles whose hit words did not fit (> kMaskWords non-empty words, extremely dense clumps) are skipped by the
// main kernel and handled here: walk all 27 cells with the exact predicate, like the variant-0 kernel.
template <bool FUSED>
__global__ void __launch_bounds__(128) k_forces_overflow(const float *__restrict__ xs, const float *__restrict__ ys,
                                                         const float *__restrict__ zs, const float4 *__restrict__ fdat,
                                                         const float4 *__restrict__ dp, const int *__restrict__ ovf,
                                                         const int *__restrict__ key, const int *__restrict__ cell_start,
                                                         float4 *__restrict__ acc, const float4 *__restrict__ pos_s,
                                                         float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                                                         int *__restrict__ far_movers, int i0, int n,
                                                         const __grid_constant__ Params P) {
    // One WARP per overflow particle (the list the density pass recorded is normally empty or short): the lanes
    // stride over the candidates of each row, and the partial sums are combined with a fixed-order butterfly, so
    // the result is still a pure function of the state.
    const int count = ovf[0];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < count; t += warps) {
        const int i = ovf[1 + t];
        if (i < i0 || i >= n) continue;  // slab mode: outside the sub-range of this launch (uniform per warp)
        float4 pi, vi;
        ld256(fdat + 2 * (size_t)i, pi, vi);
        ForceSum f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for_each_row(__ldg(key + i), cell_start, P, [&](int a, int b) {
            for (int j = a + lane; j < b; j += 32) {
                const float r2 = r2_exact(pi.x - __ldg(xs + j), pi.y - __ldg(ys + j), pi.z - __ldg(zs + j));
                if (P.h2 - r2 >= 0.0f && j != i) {
                    float4 pj, vj;
                    ld256(fdat + 2 * (size_t)j, pj, vj);
                    pair_term(f, pi, vi, pj, vj, false, P);
                }
            }
        });
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            f.px += __shfl_xor_sync(0xffffffffu, f.px, d);
            f.py += __shfl_xor_sync(0xffffffffu, f.py, d);
            f.pz += __shfl_xor_sync(0xffffffffu, f.pz, d);
            f.vx += __shfl_xor_sync(0xffffffffu, f.vx, d);
            f.vy += __shfl_xor_sync(0xffffffffu, f.vy, d);
            f.vz += __shfl_xor_sync(0xffffffffu, f.vz, d);
        }
        if (lane == 0) force_epilogue<FUSED>(i, f, pi, vi, dp, pos_s, acc, pos_out, vel_out, key, far_movers, P);
    }
}

__device__ __forceinline__ int bfind(unsigned v) {  // index of the most significant set bit (FLO.U32)
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <bool FUSED>
__global__ void __launch_bounds__(128) k_forces_mask(const float4 *__restrict__ fdat, const float4 *__restrict__ dp,
                                                     const uint2 *__restrict__ mask, const int *__restrict__ nb_words,
                                                     float4 *__restrict__ acc, const float4 *__restrict__ pos_s,
                                                     float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                                                     const int *__restrict__ key, int *__restrict__ far_movers, int i0,
                                                     int n, const __grid_constant__ Params P) {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;  // particles [i0, n): slab mode runs sub-ranges
    if (i >= n) return;
    const int nw = __ldg(nb_words + i);
    if (nw > kMaskWords) return;  // k_forces_overflow's job
    float4 pi, vi;
    ld256(fdat + 2 * (size_t)i, pi, vi);
    ForceSum f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // Flat walk: every lane consumes its own stream of set bits; a lane that runs out of bits takes its next word
    // (fetched one ahead), so the warp runs for max(hits) iterations, not the sum of per-word maxima.  The refill is
    // BRANCH-FREE: with ~21 active lanes and one refill per 6.7 hits some lane needs one in almost every iteration, and
    // a divergent 13-instruction branch for three lanes costs the warp more issue slots than eight predicated
    // instructions for everybody (52 -> 47 issue slots per iteration, -2 % time; the kernel is L1-bound).
    // nx == 0 marks "no word left", so the loop simply ends on m == 0 (stored words are never empty).
    const uint2 *wp = mask + ((size_t)(i >> 5) * kMaskWords) * 32 + (i & 31);
    unsigned m = 0u, nx = 0u, ny = 0u;
    int j31 = 0, left = nw - 1;  // words behind the current one
    if (nw > 0) {
        const uint2 t = __ldg(wp);
        m = t.x;
        j31 = (int)t.y;
    }
    if (left > 0) {
        wp += 32;
        const uint2 t = __ldg(wp);
        nx = t.x;
        ny = t.y;
    }
    static_assert(sizeof(uint2) * 32 == 256, "the refill below steps one warp-transposed word row (256 bytes)");
    while (m != 0u) {
        const int msb = bfind(m);
        m ^= 1u << msb;
        const int j = j31 - msb;
        float4 pj, vj;
        ld256(fdat + 2 * (size_t)j, pj, vj);
        pair_term(f, pi, vi, pj, vj, j == i, P);
        asm volatile(
            "{\n\t"
            ".reg .pred p, q;\n\t"
            "setp.eq.u32 p, %0, 0;\n\t"          // this lane's word ran empty:
            "@p mov.u32 %0, %2;\n\t"             //   m   = next.bits
            "@p mov.u32 %1, %3;\n\t"             //   j31 = next.j0 + 31
            "@p mov.u32 %2, 0;\n\t"              //   next = none ...
            "@p add.s32 %4, %4, -1;\n\t"
            "setp.gt.and.s32 q, %4, 0, p;\n\t"   //   ... unless another word is left: fetch it
            "@q ld.global.nc.v2.u32 {%2, %3}, [%5+256];\n\t"
            "@q add.u64 %5, %5, 256;\n\t"
            "}"
            : "+r"(m), "+r"(j31), "+r"(nx), "+r"(ny), "+r"(left), "+l"(wp));
    }
    force_epilogue<FUSED>(i, f, pi, vi, dp, pos_s, acc, pos_out, vel_out, key, far_movers, P);
}

// ================================================================= validation tap: the stored hit words as id lists
// Decodes exactly what k_forces_mask consumes — the words {bits, j0 + 31} the density pass stored — into neighbour id
// lists (self included, like the reference's predicate).  Particles whose words overflowed are enumerated the way
// k_forces_overflow walks them.  lists[offsets_by_id[id] ...] receives the ids in walk order (the host sorts).
__global__ void __launch_bounds__(128) k_mask_lists(const float4 *__restrict__ pos_s, const float *__restrict__ xs,
                                                    const float *__restrict__ ys, const float *__restrict__ zs,
                                                    const uint2 *__restrict__ mask, const int *__restrict__ nb_words,
                                                    const int *__restrict__ key, const int *__restrict__ cell_start,
                                                    const long long *__restrict__ offsets_by_id, int *__restrict__ lists,
                                                    int *__restrict__ counts_by_id, int n, const __grid_constant__ Params P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned id = __float_as_uint(__ldg(&pos_s[i].w));
    if (id >= (unsigned)n) return;
    long long w = lists ? offsets_by_id[id] : 0;
    int cnt = 0;
    const int nw = __ldg(nb_words + i);
    if (nw <= kMaskWords) {
        const uint2 *const wbase = mask + ((size_t)(i >> 5) * kMaskWords) * 32 + (i & 31);
        for (int k = 0; k < nw; ++k) {
            const uint2 word = __ldg(wbase + k * 32);
            unsigned m = word.x;
            while (m) {
                const int msb = 31 - __clz(m);
                m &= ~(1u << msb);
                const int j = (int)word.y - msb;
                if (lists) lists[w++] = __float_as_int(__ldg(&pos_s[j].w));
                ++cnt;
            }
        }
    } else {
        const float px = __ldg(xs + i), py = __ldg(ys + i), pz = __ldg(zs + i);
        for_each_row(__ldg(key + i), cell_start, P, [&](int a, int b) {
            for (int j = a; j < b; ++j) {
                const float r2 = r2_exact(px - __ldg(xs + j), py - __ldg(ys + j), pz - __ldg(zs + j));
                if (P.h2 - r2 >= 0.0f) {
                    if (lists) lists[w++] = __float_as_int(__ldg(&pos_s[j].w));
                    ++cnt;
                }
            }
        });
    }
    if (counts_by_id) counts_by_id[id] = cnt;
}

void launch_mask_lists(const NbBuffers &nb, const float4 *pos_s, const int *key_s, const int *cell_start,
                       const long long *offsets_by_id, int *lists, int *counts_by_id, int n, const Params &P, cudaStream_t st) {
    if (n <= 0) return;
    k_mask_lists<<<(n + 127) / 128, 128, 0, st>>>(pos_s, nb.xs, nb.ys, nb.zs, nb.mask, nb.words, key_s, cell_start, offsets_by_id,
                                                  lists, counts_by_id, n, P);
}

void launch_forces_mask(const NbBuffers &nb, const float4 *dp, const int *nb_count, const int *key_s, const int *cell_start,
                        float4 *acc, int i0, int i1, const Params &P, cudaStream_t st, const float4 *pos_s, float4 *pos_out,
                        float4 *vel_out, int *far_movers) {
    if (i1 <= i0) return;
    (void)nb_count;
    const int grid = (i1 - i0 + 127) / 128;
    if (pos_out) {  // fused with walls + integration
        k_forces_mask<true><<<grid, 128, 0, st>>>(nb.fdat, dp, nb.mask, nb.words, acc, pos_s, pos_out, vel_out, key_s,
                                                  far_movers, i0, i1, P);
        k_forces_overflow<true><<<148 * 4, 128, 0, st>>>(nb.xs, nb.ys, nb.zs, nb.fdat, dp, nb.ovf, key_s, cell_start, acc,
                                                         pos_s, pos_out, vel_out, far_movers, i0, i1, P);
    } else {
        k_forces_mask<false><<<grid, 128, 0, st>>>(nb.fdat, dp, nb.mask, nb.words, acc, nullptr, nullptr, nullptr, nullptr,
                                                   nullptr, i0, i1, P);
        k_forces_overflow<false><<<148 * 4, 128, 0, st>>>(nb.xs, nb.ys, nb.zs, nb.fdat, dp, nb.ovf, key_s, cell_start, acc,
                                                          nullptr, nullptr, nullptr, nullptr, i0, i1, P);
    }
}

}  // namespace sph
