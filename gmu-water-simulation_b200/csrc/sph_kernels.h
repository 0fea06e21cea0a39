// sph_kernels.h — host-side launchers of the sm_100a kernels (internal to libsph_cuda.so).
#pragma once

#include "sph_device.cuh"

namespace sph {

constexpr int kScanVec = 4;                    // int4 vectors per thread in the scan
constexpr int kScanTile = 512 * 4 * kScanVec;  // cells per scan tile (512 threads x 4 x int4 = 8192): short look-back chains
constexpr int kMaskWords = 40;   // stored (non-empty) hit words per particle; more -> overflow path (warp per particle)
constexpr int kSoaPad = 64;      // far-away sentinel entries after the last particle of xs/ys/zs

// Extra arrays of the production neighbour passes (sph_neighbours_v2.cu)
struct NbBuffers {
    float *xs, *ys, *zs;  // [cap + kSoaPad] canonical-order positions, split SoA (16-byte loads = 4 candidates)
    float4 *fdat;         // [2*cap] {x,y,z,p/rho^2 | vx,vy,vz,1/rho}: one 256-bit gather per neighbour
    uint2 *mask;          // [ceil(cap/32)*kMaskWords*32] hit words {bits, first candidate + 31}, warp-transposed:
                          // word w of lane l of warp q at (q*kMaskWords + w)*32 + l
    int *words;           // [cap] number of non-empty hit words of each particle (> kMaskWords = overflow)
    int *ovf;             // [cap + 1] ovf[0] = number of overflow particles of this step, ovf[1..] = their indices
};

struct GridBuffers {
    int *key_a;                       // [cap]   cell id in input order
    int *off_a;                       // [cap]   arrival offset inside the cell (atomic)
    int *bucket_src;                  // [cap]   input index per bucket slot
    int *bucket_id;                   // [cap]   particle id per bucket slot
    int *key_s;                       // [cap]   sort key (reference cell id refined by the x bin) in canonical order
    int *count;                       // [cells_padded] per-cell histogram (zeroed by the scan)
    int *cell_start;                  // [cells_padded] exclusive scan; [n_cells] = n
    unsigned long long *scan_status;  // [tiles + 1]: tile look-back words, last = tile ticket
    int n_scan_items;                 // n_cells * xb + 1
    int n_tiles;
    int scan_tile0, scan_tiles;       // the tiles launch_scan covers: all of them, or (slab mode) those of the layers in use
};

// grid build: keys + histogram, exclusive scan, bucket, rank-by-id + SoA reorder
void launch_cell_key_hist(const float4 *pos_a, int n, const GridBuffers &g, const Params &P, cudaStream_t st);
void launch_scan(const GridBuffers &g, cudaStream_t st);
void launch_bucket(const float4 *pos_a, int n, const GridBuffers &g, cudaStream_t st);
void launch_rank_scatter(const float4 *pos_a, const float4 *vel_a, float4 *pos_s, float4 *vel_s, int n,
                         const GridBuffers &g, const NbBuffers &nb, cudaStream_t st);

// neighbour passes on the canonical order
void launch_density(const float4 *pos_s, const int *key_s, const int *cell_start, float4 *dp, int *nb_count, int n,
                    const Params &P, cudaStream_t st);
void launch_forces(const float4 *pos_s, const float4 *vel_s, const float4 *dp, const int *key_s, const int *cell_start,
                   float4 *acc, int n, const Params &P, cudaStream_t st);
void launch_density_mask(const NbBuffers &nb, const float4 *vel_s, const int *key_s, const int *cell_start, float4 *dp,
                         int *nb_count, int n, const Params &P, cudaStream_t st, bool reset_overflow_list);
// [i0, i1) = index range of the canonical order to process (the whole array outside slab mode)
// pos_out != NULL: fused with the wall term + integration (new state written to pos_out/vel_out, pos_s supplies the ids)
void launch_forces_mask(const NbBuffers &nb, const float4 *dp, const int *nb_count, const int *key_s, const int *cell_start,
                        float4 *acc, int i0, int i1, const Params &P, cudaStream_t st, const float4 *pos_s = nullptr,
                        float4 *pos_out = nullptr, float4 *vel_out = nullptr, int *far_movers = nullptr);
// far_movers (may be NULL): slab mode counter of particles that crossed more than 2 z-layers in this step
void launch_integrate_collide(const float4 *pos_s, const float4 *vel_s, float4 *acc, float4 *pos_out, float4 *vel_out,
                              int i0, int i1, const int *key_s, int *far_movers, const Params &P, cudaStream_t st);

// all-pairs validation kernels (CGPUBruteParticleSimulator semantics)
void launch_brute_density(const float4 *pos, float4 *dp, int *nb_count, int n, const Params &P, cudaStream_t st);
void launch_brute_forces(const float4 *pos, const float4 *vel, const float4 *dp, float4 *acc, int n, const Params &P,
                         cudaStream_t st);

// AoS <-> SoA at the boundary, taps
struct ParticleAoS {  // == sph_particle
    float position[4], velocity[4], acceleration[4];
    int grid_position[4];
    float density, pressure;
    unsigned id, cell_id;
};
// id_error (may be NULL): set to 1 when a record's id is >= id_limit
void launch_aos_to_soa(const ParticleAoS *aos, float4 *pos, float4 *vel, int n, unsigned id_limit, int *id_error,
                       cudaStream_t st);
// id_base >= 0: record of particle `id` goes to aos[id - id_base] (ids outside [id_base, id_base+id_count) are skipped);
// id_base < 0 : compact mode, particle at index i goes to aos[i] (slab mode: owned sub-range, order = canonical order)
void launch_soa_to_aos(const float4 *pos, const float4 *vel, const float4 *acc, const float4 *dp, const int *key,
                       ParticleAoS *aos_by_id, int id_base, int id_count, int n, const Params &P, cudaStream_t st);
// fountain emission on the device: n_new template records appended, ids id_base .. id_base + n_new - 1
void launch_emit(const float4 *tpl_pos, const float4 *tpl_vel, int n_new, float4 *pos_out, float4 *vel_out, unsigned id_base,
                 cudaStream_t st);
// slab mode: append the owned particles lying in the 2+2 layers around each face to the send buffers
void launch_gather4(const int *src, size_t i0, size_t i1, size_t i2, size_t i3, int *dst, cudaStream_t st);
void launch_slab_pack(const float4 *pos, const float4 *vel, int n, int z_lo_below, int z_hi_from, float4 *down_pos,
                      float4 *down_vel, float4 *up_pos, float4 *up_vel, int *counters, int cap_face, const Params &P,
                      cudaStream_t st);
void launch_scatter_by_id_i32(const float4 *pos, const int *src, int *dst_by_id, int n, cudaStream_t st);
void launch_scatter_cell_ids(const float4 *pos, const int *key, int *dst_by_id, int n, const Params &P, cudaStream_t st);
void launch_coarse_cell_start(const int *cell_start, int *out, int n_cells, int xb, cudaStream_t st);
void launch_cell_id_permutation(const float4 *pos, const int *key, const int *cell_start, unsigned *out, int n,
                                const Params &P, cudaStream_t st);
void launch_scatter_dpa_by_id(const float4 *pos, const float4 *dp, const float4 *acc, float *rho, float *p, float *acc3,
                              int n, cudaStream_t st);
void launch_extract_ids(const float4 *pos, unsigned *ids, int n, cudaStream_t st);
// neighbour lists: pass 1 counts (by id), pass 2 fills lists at offsets[id] (unsorted; host sorts each list)
void launch_neighbour_lists(const float4 *pos_s, const int *key_s, const int *cell_start, const long long *offsets_by_id,
                            int *lists, int n, const Params &P, cudaStream_t st);
// the stored hit words (what the force pass consumes) decoded into id lists; lists == NULL: counts only
void launch_mask_lists(const NbBuffers &nb, const float4 *pos_s, const int *key_s, const int *cell_start,
                       const long long *offsets_by_id, int *lists, int *counts_by_id, int n, const Params &P, cudaStream_t st);
void launch_flush_l2(float4 *buf, size_t count, cudaStream_t st);
void launch_stats(const float4 *pos, const float4 *vel, int n, double *out8, cudaStream_t st);
void launch_sum_i32(const int *src, int n, unsigned long long *out, cudaStream_t st);
// histogram of floor((y - lo) * inv_width) over kHistBins bins; `clamp` puts out-of-range values into the edge bins,
// otherwise they are not counted (refinement pass inside one bin)
constexpr int kHistBins = 4096;
void launch_minmax_y(const float4 *pos, int n, unsigned *out2, cudaStream_t st);  // ordered-uint bits of min / max finite y
void launch_hist_y(const float4 *pos, int n, double lo, double inv_width, int clamp, unsigned *hist, cudaStream_t st);

}  // namespace sph
