// sph_grid.cu — device uniform grid: cell keys -> counting sort -> SoA reorder (sm_100a).
//
// Replaces CCPUParticleSimulator::updateGrid (src/CCPUParticleSimulator.cpp:32-91) and the OpenCL
// path's histogram + Blelloch scan + *host* std::sort (src/CGPUParticleSimulator.cpp:57-139).
// Nothing leaves the device.  The output order is canonical: stable by (cell_id, particle id), so
// the permutation and every later summation order are pure functions of the particle state.
//
//   K1 cell_key_hist : key[i] = cell(pos[i]) in fp64, off[i] = atomicAdd(count[key], 1)
//   K2 scan          : single-pass decoupled look-back exclusive scan of count -> cell_start (8192 cells per tile),
//                      zeroes count for the next step
//   K3 bucket        : slot = cell_start[key] + off  ->  bucket_src/bucket_id   (arrival order)
//   K4 rank_scatter  : rank inside the cell by particle id, gather pos/vel into canonical order; also re-arms the
//                      scan's look-back words and the overflow list for the next use (no memset nodes in the step)
#include "sph_kernels.h"

namespace sph {

// ---------------------------------------------------------------- K1
__global__ void __launch_bounds__(256) k_cell_key_hist(const float4 *__restrict__ pos, int n, int *__restrict__ key,
                                                       int *__restrict__ off, int *__restrict__ count,
                                                       const __grid_constant__ Params P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pos + i);
    const int k = cell_key(p, P);
    key[i] = k;
    off[i] = atomicAdd(count + k, 1);
}

void launch_cell_key_hist(const float4 *pos_a, int n, const GridBuffers &g, const Params &P, cudaStream_t st) {
    if (n <= 0) return;
    k_cell_key_hist<<<(n + 255) / 256, 256, 0, st>>>(pos_a, n, g.key_a, g.off_a, g.count, P);
}

// ---------------------------------------------------------------- K2
// Status word: bits 63..62 = flag (0 empty, 1 tile aggregate, 2 inclusive prefix), low 32 bits = value.
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagMask = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(512) k_scan(int *__restrict__ count, int *__restrict__ cell_start, int n_items,
                                              unsigned long long *__restrict__ status, int n_tiles, int tile0) {
    __shared__ int s_tile;
    __shared__ int s_warp[16];
    __shared__ int s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int)atomicAdd(status + n_tiles, 1ull);  // ticket: tiles start in order
    __syncthreads();
    const int tile = s_tile;
    // tile0: first tile of the launch (slab mode scans only the layers that can hold particles; the bins below it are
    // empty, so the running prefix starts at 0 there as well)
    const int base = (tile0 + tile) * kScanTile + tid * (4 * kScanVec);

    // buffers are padded to a multiple of the tile and the padding stays zero
    int4 c[kScanVec];
    int t_sum = 0;
#pragma unroll
    for (int v = 0; v < kScanVec; ++v) {
        c[v] = *reinterpret_cast<const int4 *>(count + base + 4 * v);
        *reinterpret_cast<int4 *>(count + base + 4 * v) = make_int4(0, 0, 0, 0);
        t_sum += c[v].x + c[v].y + c[v].z + c[v].w;
    }

    int incl = t_sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 16 ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += v;
        }
        if (lane < 16) s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int warp_excl = warp ? s_warp[warp - 1] : 0;
    const int tile_total = s_warp[15];
    const int thread_excl = warp_excl + incl - t_sum;

    // decoupled look-back by warp 0
    if (warp == 0) {
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) st_status(status, kFlagPrefix | (unsigned)tile_total);
        } else {
            if (lane == 0) st_status(status + tile, kFlagAgg | (unsigned)tile_total);
            int look = tile - 1;
            while (true) {
                const int t = look - lane;
                unsigned long long w = kFlagPrefix;  // tiles below 0 count as a zero prefix
                if (t >= 0) {
                    do { w = ld_status(status + t); } while ((w & kFlagMask) == 0);
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (w & kFlagMask) == kFlagPrefix);
                // add lanes up to and including the first one that carries an inclusive prefix
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;
                int v = (lane <= first) ? (int)(unsigned)(w & 0xffffffffull) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                prefix += v;
                if (has_prefix) break;
                look -= 32;
            }
            if (lane == 0) st_status(status + tile, kFlagPrefix | (unsigned)(prefix + tile_total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    int e = s_prefix + thread_excl;
#pragma unroll
    for (int v = 0; v < kScanVec; ++v) {
        int4 o;
        o.x = e;
        o.y = e + c[v].x;
        o.z = o.y + c[v].y;
        o.w = o.z + c[v].z;
        e = o.w + c[v].w;
        *reinterpret_cast<int4 *>(cell_start + base + 4 * v) = o;
    }
    (void)n_items;
}

void launch_scan(const GridBuffers &g, cudaStream_t st) {
    // scan_status is all zero here: zeroed at creation and again by every k_rank_scatter (which follows each scan)
    k_scan<<<g.scan_tiles, 512, 0, st>>>(g.count, g.cell_start, g.n_scan_items, g.scan_status, g.n_tiles, g.scan_tile0);
}

// ---------------------------------------------------------------- K3
__global__ void __launch_bounds__(256) k_bucket(const float4 *__restrict__ pos, int n, const int *__restrict__ key,
                                                const int *__restrict__ off, const int *__restrict__ cell_start,
                                                int *__restrict__ bucket_src, int *__restrict__ bucket_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = __ldg(cell_start + key[i]) + off[i];
    bucket_src[slot] = i;
    bucket_id[slot] = __float_as_int(__ldg(&pos[i].w));
}

void launch_bucket(const float4 *pos_a, int n, const GridBuffers &g, cudaStream_t st) {
    if (n <= 0) return;
    k_bucket<<<(n + 255) / 256, 256, 0, st>>>(pos_a, n, g.key_a, g.off_a, g.cell_start, g.bucket_src, g.bucket_id);
}

// ---------------------------------------------------------------- K4
__global__ void __launch_bounds__(256) k_rank_scatter(const float4 *__restrict__ pos_a, const float4 *__restrict__ vel_a,
                                                      float4 *__restrict__ pos_s, float4 *__restrict__ vel_s,
                                                      int *__restrict__ key_s, int n, const int *__restrict__ key_a,
                                                      const int *__restrict__ cell_start,
                                                      const int *__restrict__ bucket_src,
                                                      const int *__restrict__ bucket_id, float *__restrict__ xs,
                                                      float *__restrict__ ys, float *__restrict__ zs,
                                                      unsigned long long *__restrict__ scan_status, int n_status,
                                                      int *__restrict__ ovf) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    // housekeeping that used to be two memset nodes: the scan of this grid build is complete (stream order), and
    // the overflow list of the previous density pass has been consumed
    if (s < n_status) scan_status[s] = 0ull;
    if (s == 0 && ovf) ovf[0] = 0;
    if (s >= n) {
        // sentinel tail: the 4-wide candidate loads may run up to 3 entries past the last particle
        if (xs && s < n + kSoaPad) xs[s] = ys[s] = zs[s] = 1.0e18f;
        return;
    }
    const int src = bucket_src[s];
    const int my_id = bucket_id[s];
    const int k = __ldg(key_a + src);
    const int a = __ldg(cell_start + k), b = __ldg(cell_start + k + 1);
    int rank = 0;
    for (int t = a; t < b; ++t) rank += (__ldg(bucket_id + t) < my_id);
    const int dst = a + rank;
    const float4 p = __ldg(pos_a + src);
    pos_s[dst] = p;
    vel_s[dst] = __ldg(vel_a + src);
    key_s[dst] = k;
    if (xs) {
        // A particle with a non-finite coordinate can be nobody's neighbour in the reference (every comparison
        // with NaN is false).  The bitmask pass takes the hit from the sign bit of h2 - r2, so such a particle is
        // given the far-away sentinel in the candidate arrays instead; pos_s keeps its real value.
        const bool ok = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
        xs[dst] = ok ? p.x : 1.0e18f;
        ys[dst] = ok ? p.y : 1.0e18f;
        zs[dst] = ok ? p.z : 1.0e18f;
    }
}

void launch_rank_scatter(const float4 *pos_a, const float4 *vel_a, float4 *pos_s, float4 *vel_s, int n,
                         const GridBuffers &g, const NbBuffers &nb, cudaStream_t st) {
    if (n < 0) n = 0;
    const int threads = (n + kSoaPad > g.n_tiles + 1) ? n + kSoaPad : g.n_tiles + 1;
    k_rank_scatter<<<(threads + 255) / 256, 256, 0, st>>>(pos_a, vel_a, pos_s, vel_s, g.key_s, n, g.key_a, g.cell_start,
                                                          g.bucket_src, g.bucket_id, nb.xs, nb.ys, nb.zs, g.scan_status,
                                                          g.n_tiles + 1, nb.ovf);
}

}  // namespace sph
