// sph_capi.cu — the C ABI of include/sph_cuda.h: context, buffers, phase sequencing, CUDA-graph step,
// event timing and error strings.  Replaces CLWrapper (src/CLWrapper.cpp) and CLPlatforms
// (src/CLPlatforms.cpp) for the SPH path.  No CPU fallback anywhere: without a usable CUDA device
// every entry point returns SPH_ERR_CUDA.
#include "../../include/sph_cuda.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved lazily with dlopen (slab mode), never linked

#include "sph_kernels.h"

using namespace sph;

static_assert(sizeof(sph_particle) == 80, "sph_particle must match CParticle::Physics (80 B)");
static_assert(sizeof(ParticleAoS) == 80, "ParticleAoS must match sph_particle");
static_assert(sizeof(sph_wall) == 32, "sph_wall must match sWall (32 B)");

namespace {
thread_local std::string g_last_error;  // for failures without a context
}

struct sph_context {
    sph_config cfg;
    Params P;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t cap = 0, n = 0;
    // A = authoritative state (pos/vel), S = canonical-order snapshot the neighbour passes read
    float4 *pos_a = nullptr, *vel_a = nullptr, *pos_s = nullptr, *vel_s = nullptr;
    float4 *dp = nullptr, *acc = nullptr;
    int *nb_count = nullptr;
    GridBuffers g{};
    NbBuffers nb{};
    size_t cells_padded = 0;
    bool s_valid = false;     // S holds the pre-integration snapshot the aux arrays are aligned with
    bool grid_valid = false;  // S is in canonical order, key_s / cell_start valid
    bool a_aligned = false;   // A is in the same order as S (true after integrate)
    bool density_valid = false, forces_valid = false;
    bool ovf_clean = false;  // the overflow list is empty (k_rank_scatter just ran, no density pass since)
    ParticleAoS *d_stage = nullptr;
    size_t stage_cap = 0;
    int *d_tmp_i32 = nullptr;  // [cap] scratch for by-id taps
    float *d_tmp_f32 = nullptr;  // [5*cap]
    double *d_stats = nullptr;
    int *d_id_error = nullptr;  // set by the upload kernel when a record's id is >= max_particles (ABI precondition)
    unsigned *d_hist = nullptr;  // [kHistBins] fill-height histogram (sph_fill_height_percentile)
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_n = 0;
    uint32_t last_step_n = 0xffffffffu;
    int opt_neighbour_variant = 1, opt_use_graph = 1, opt_count_neighbours = 1, opt_fuse_integrate = 1;
    uint64_t kernel_launches = 0, graph_launches = 0, steps = 0;
    std::string err;
    void *pinned_ptr = nullptr;
    float4 *d_mesh_planes = nullptr;  // sph_set_collision_faces
    // asynchronous read-back (sph_download_particles_async): snapshot on `stream`, PCIe copy on `copy_stream`
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copy = nullptr;
    bool copy_pending = false;
    uint32_t copy_n = 0;
    // bench hygiene: evict L2 between timed steps (option "flush_l2") and time each step separately
    int opt_flush_l2 = 0;
    float4 *d_flush = nullptr;
    size_t flush_count = 0;
    std::vector<cudaEvent_t> step_events;
    // device-side state snapshots (sph_state_save / sph_state_restore): A arrays in their current order
    float4 *snap_pos[2] = {nullptr, nullptr}, *snap_vel[2] = {nullptr, nullptr};
    uint32_t snap_n[2] = {0, 0};
    // device-side fountain emitter (sph_set_emitter): templates of the records one step appends
    float4 *d_emit_pos = nullptr, *d_emit_vel = nullptr;
    uint32_t emit_templates = 0, emit_group = 0, emit_max = 0;
    // slab mode (multi-GPU): this rank's z-range, exchange buffers and NCCL communicator
    struct Slab *slab = nullptr;
    uint32_t in_off = 0;  // the next grid build reads A[in_off, in_off + n)
};

// ---- loop-back transport ----------------------------------------------------------------------------
// The slab exchange needs two operations from its transport: "tell both neighbours how many particles are coming"
// and "deliver the face buffers into the neighbour's tail".  Between processes that is NCCL send/recv (below).
// The loop-back transport provides the same two operations between `world` contexts that live in ONE process (on
// one device or several): every rank runs in its own host thread exactly like a rank process would, counts travel
// through a mailbox in host memory and payloads as device-to-device copies ordered by CUDA events.  It exists so
// that the k-slab == 1-GPU equivalence can be checked bit for bit on a one-GPU box; everything above
// the transport (packing, ghost layers, ownership, overlap schedule) is the code the NCCL runs use.
struct LocalEdge {  // one direction of one face: src -> dst
    std::mutex m;
    std::condition_variable cv;
    uint64_t posted = 0, consumed = 0;
    int count = 0;
    int aux[4] = {0, 0, 0, 0};     // the sender's {load, z0, z1, can_grow} (face rebalancing, see Slab)
    const float4 *pos = nullptr, *vel = nullptr;
    cudaEvent_t packed = nullptr;  // sender: face buffers complete
    cudaEvent_t copied = nullptr;  // receiver: face buffers read, may be overwritten
    bool aborted = false;
};
struct LocalComm {
    int world = 0;
    std::unique_ptr<LocalEdge[]> up, down;  // up[r]: r -> r + 1, down[r]: r -> r - 1
    std::atomic<int> refs{0};
};
constexpr char kLocalMagic[8] = {'S', 'P', 'H', 'L', 'O', 'O', 'P', '1'};

// One rank of a 1-D slab decomposition along z (the slowest cell axis, so a slab is a contiguous key range
// and, after the sort, a contiguous particle range).
struct Slab {
    int rank = 0, world = 1;
    int z0 = 0, z1 = 0;        // owned global layers [z0, z1)
    int z_base = 0, rz_local = 0;
    uint32_t n_own = 0;        // owned particles: A[in_off, in_off + n_own)
    int cap_face = 0;          // capacity (particles) of each face buffer
    float4 *down_pos = nullptr, *down_vel = nullptr, *up_pos = nullptr, *up_vel = nullptr;
    int *d_counters = nullptr;  // [0] to lower, [1] to upper, [2] from lower, [3] from upper
    ncclComm_t comm = nullptr;
    LocalComm *local = nullptr;  // loop-back transport instead of NCCL (all ranks in this process)
    uint64_t sent_particles = 0, exchanges = 0;
    // overlapped exchange: pack + NCCL run on their own stream while the interior of the slab is still computing
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ranges = nullptr, ev_boundary = nullptr, ev_counts = nullptr, ev_comm = nullptr;
    bool have_ghosts = false;  // A already holds this step's ghosts/migrants behind the owned particles
    int *h_pinned = nullptr;   // 32 ints of pinned host memory: [0..3] counts, [4..7] layer starts, [8..11] aux out, [12..19] aux in
    int opt_overlap = 1;
    // ---- moving faces (load balance).  Equal layer counts are not equal work: the neighbour count per particle varies
    // along the tank, and every step runs at the pace of the slowest slab.  With every exchange a rank therefore also
    // sends {busy time of its last complete step, z0, z1, may-grow}; both ranks at a face apply the same rule to the same
    // two records (face_shift) and move the face by at most one layer per step towards the busier side.  Nothing else
    // has to change: the rank that gives a layer away packs it as "beyond the face" (it integrated it, so the data is
    // current) and keeps its own copy as a ghost layer; ownership is decided by the cell layer at the next sort.  The
    // local grid is allocated with shift_max spare layers on either side, so z_base never moves.
    int z0_init = 0, z1_init = 0, shift_max = 0;
    int opt_rebalance = 1;     // 0 = static faces, 1 = by load, 2 = test pattern (faces oscillate deterministically)
    int aux_sent[4] = {0, 0, 0, 0};      // what this rank sent with the last exchange
    int aux_recv[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};  // what it received from below [0] / above [1]
    bool aux_valid = false;
    int *d_aux = nullptr;      // device: [0..3] out, [4..7] in from below, [8..11] in from above
    cudaEvent_t ev_busy[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [step parity][begin, end] of the compute span
    int busy_recorded[2] = {0, 0};
    int load_us = 0;
    uint64_t face_moves = 0;
};

// The rule both ranks at a face evaluate on the same two records: -1 = the face moves down (the lower rank gives its
// top layer to the upper rank), +1 = up, 0 = stays.  lo / hi = {load, z0, z1, can_grow} of the lower / upper rank;
// shift = current offset of the face from the initial plan.
int face_shift(const int lo[4], const int hi[4], int shift, int shift_max, int mode, uint64_t exchange, int face) {
    constexpr int kMinLayers = 10;
    const int layers_lo = lo[2] - lo[1], layers_hi = hi[2] - hi[1];
    int want = 0;
    if (mode == 2) {  // test pattern: three exchanges up, three down, phase-shifted per face
        want = ((exchange + (uint64_t)face) % 6) < 3 ? +1 : -1;
    } else if (mode == 1 && lo[0] > 0 && hi[0] > 0) {
        if ((long long)lo[0] * 1000 > (long long)hi[0] * 1004) want = -1;     // lower rank is busier: it gives a layer away
        else if ((long long)hi[0] * 1000 > (long long)lo[0] * 1004) want = +1;
    }
    if (want < 0 && (layers_lo <= kMinLayers || shift <= -shift_max || !hi[3])) want = 0;
    if (want > 0 && (layers_hi <= kMinLayers || shift >= shift_max || !lo[3])) want = 0;
    return want;
}

namespace {

int fail(sph_context *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(ctx, call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, SPH_ERR_CUDA,                                                                    \
                        std::string("CUDA::") + cudaGetErrorName(e_) + ": " + cudaGetErrorString(e_) + " | " #call); \
    } while (0)

#define REQUIRE(ctx, cond, code, msg) \
    do {                              \
        if (!(cond)) return fail(ctx, code, msg); \
    } while (0)

int check_launch(sph_context *ctx, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SPH_ERR_CUDA, std::string("CUDA::") + cudaGetErrorName(e) + " after " + what);
    return SPH_OK;
}

Params make_params(const sph_config &c) {
    Params P{};
    P.rx = c.grid_res[0];
    P.ry = c.grid_res[1];
    P.rz = c.grid_res[2];
    P.n_cells = P.rx * P.ry * P.rz;
    P.xb = 1;
    P.h_win = (double)c.h * (1.0 + 1e-6);
    P.rz_global = P.rz;
    P.z_base = 0;
    P.own_z0 = P.next_z0 = 0;
    P.own_z1 = P.next_z1 = P.rz;
    P.kz_lo = 0;
    P.kz_hi = P.rz;
    P.dens_key_lo = 0;
    P.dens_key_hi = 0x7fffffff;
    // src/CCPUParticleSimulator.cpp:46-48: m_boxSize.x() / 2.0 and CParticle::h widen to double
    P.hbx = (double)c.box[0] / 2.0;
    P.hby = (double)c.box[1] / 2.0;
    P.hbz = (double)c.box[2] / 2.0;
    P.h_d = (double)c.h;
    P.h = c.h;
    P.h2 = c.h * c.h;  // fp32 product, as CParticle::h * CParticle::h
    P.hbx_f = (float)P.hbx;
    P.hby_f = (float)P.hby;
    P.hbz_f = (float)P.hbz;
    P.neg_zero = -0.0f;
    P.prune_margin = 1e-3f * c.h + 8.0f * 1.2e-7f * std::max(c.box[0], std::max(c.box[1], c.box[2]));
    P.dt = c.dt;
    P.mass = c.mass;
    P.viscosity = c.viscosity;
    P.gas_stiffness = c.gas_stiffness;
    P.rest_density = c.rest_density;
    // fp64 coefficients of the CPU path (src/CCPUParticleSimulator.cpp:11,19,26), rounded once
    const double pi = 3.14159265358979323846;
    P.poly6_f = (float)(315.0 / (64.0 * pi * std::pow((double)c.h, 9)));
    P.spiky_f = (float)(-45.0 / (pi * std::pow((double)c.h, 6)));
    P.visc_f = (float)(45.0 / (pi * std::pow((double)c.h, 6)));
    P.gx = c.gravity[0];
    P.gy = c.gravity[1];
    P.gz = c.gravity[2];
    P.wall_k_f = (float)c.wall_k;  // narrowed where it meets a QVector3D (float overloads only)
    P.wall_damping_d = c.wall_damping;
    P.wall_skin_d = c.wall_skin;
    P.wall_count = c.wall_count;
    for (int w = 0; w < 6; ++w) {
        P.walls[w] = {c.walls[w].normal[0],   c.walls[w].normal[1],   c.walls[w].normal[2],
                      c.walls[w].position[0], c.walls[w].position[1], c.walls[w].position[2]};
    }
    return P;
}

template <typename T>
cudaError_t dalloc(T **p, size_t count) {
    return cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(count, 1) * sizeof(T));
}

struct PhaseTimer {
    sph_context *c;
    double *ms;
    PhaseTimer(sph_context *ctx, double *out) : c(ctx), ms(out) {
        if (ms) cudaEventRecord(c->ev0, c->stream);
    }
    int finish() {
        if (!ms) return SPH_OK;
        cudaEventRecord(c->ev1, c->stream);
        cudaError_t e = cudaEventSynchronize(c->ev1);
        if (e != cudaSuccess) return fail(c, SPH_ERR_CUDA, std::string("CUDA::") + cudaGetErrorName(e) + " in phase");
        float t = 0.f;
        cudaEventElapsedTime(&t, c->ev0, c->ev1);
        *ms = (double)t;
        return SPH_OK;
    }
};

// ---- the pipeline pieces (enqueue only) -------------------------------------------------------
void enqueue_grid(sph_context *c) {
    const int n = (int)c->n;
    const float4 *pa = c->pos_a + c->in_off, *va = c->vel_a + c->in_off;
    launch_cell_key_hist(pa, n, c->g, c->P, c->stream);
    launch_scan(c->g, c->stream);
    launch_bucket(pa, n, c->g, c->stream);
    launch_rank_scatter(pa, va, c->pos_s, c->vel_s, n, c->g, c->nb, c->stream);
    c->ovf_clean = true;
    c->kernel_launches += 4;
}
// variant 1 (default): bitmask passes of sph_neighbours_v2.cu; variant 0: the plain float4 walk of
// sph_neighbours.cu, also used for grids narrower than 4 cells in x (row over-scan could wrap around)
bool use_mask_passes(const sph_context *c) { return c->opt_neighbour_variant == 1 && c->P.rx >= 4; }
void enqueue_density(sph_context *c) {
    if (use_mask_passes(c))
        launch_density_mask(c->nb, c->vel_s, c->g.key_s, c->g.cell_start, c->dp, c->nb_count, (int)c->n, c->P, c->stream,
                            !c->ovf_clean),
            c->ovf_clean = false;
    else
        launch_density(c->pos_s, c->g.key_s, c->g.cell_start, c->dp, c->nb_count, (int)c->n, c->P, c->stream);
    c->kernel_launches += 1;
}
void enqueue_forces(sph_context *c) {
    if (use_mask_passes(c))
        launch_forces_mask(c->nb, c->dp, c->nb_count, c->g.key_s, c->g.cell_start, c->acc, 0, (int)c->n, c->P, c->stream),
            c->kernel_launches += 1;  // main kernel + the (normally idle) overflow kernel
    else
        launch_forces(c->pos_s, c->vel_s, c->dp, c->g.key_s, c->g.cell_start, c->acc, (int)c->n, c->P, c->stream);
    c->kernel_launches += 1;
}
void enqueue_integrate(sph_context *c) {
    launch_integrate_collide(c->pos_s, c->vel_s, c->acc, c->pos_a, c->vel_a, 0, (int)c->n, nullptr, nullptr, c->P, c->stream);
    c->kernel_launches += 1;
    c->in_off = 0;  // integrate writes A[0, n) in canonical order
}
void enqueue_step(sph_context *c) {
    enqueue_grid(c);
    enqueue_density(c);
    if (use_mask_passes(c) && c->opt_fuse_integrate) {
        // forces + walls + integration in one kernel (plus the normally idle overflow kernel)
        launch_forces_mask(c->nb, c->dp, c->nb_count, c->g.key_s, c->g.cell_start, c->acc, 0, (int)c->n, c->P, c->stream,
                           c->pos_s, c->pos_a, c->vel_a);
        c->kernel_launches += 2;
        c->in_off = 0;
    } else {
        enqueue_forces(c);
        enqueue_integrate(c);
    }
}
constexpr int kKernelsPerStep = 7;

void drop_graph(sph_context *c) {
    if (c->graph_exec) {
        cudaGraphExecDestroy(c->graph_exec);
        c->graph_exec = nullptr;
    }
    c->graph_n = 0;
}

// pos/vel arrays the taps should read: S between update_grid and integrate, A otherwise
const float4 *view_pos(const sph_context *c) { return c->a_aligned || !c->s_valid ? c->pos_a : c->pos_s; }
const float4 *view_vel(const sph_context *c) { return c->a_aligned || !c->s_valid ? c->vel_a : c->vel_s; }
// dp/acc/key_s index == view index; a new grid build invalidates last step's density / forces until they are recomputed
bool aux_aligned(const sph_context *c) { return c->s_valid; }
const float4 *aux_dp(const sph_context *c) { return (c->s_valid && c->density_valid) ? c->dp : nullptr; }
const float4 *aux_acc(const sph_context *c) { return (c->s_valid && c->forces_valid) ? c->acc : nullptr; }

// By-id read-backs index host arrays with the particle ids: they need the ids to be < max_particles (checked by the
// upload kernel) and, to fill every slot, a permutation of 0..n-1.  Called after the tap's own stream sync.
int id_precondition(sph_context *c, const char *who) {
    int flag = 0;
    cudaError_t e = cudaMemcpyAsync(&flag, c->d_id_error, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(c, SPH_ERR_CUDA, std::string("CUDA::") + cudaGetErrorName(e) + " | id check");
    if (flag)
        return fail(c, SPH_ERR_ARGUMENT, std::string(who) + ": an uploaded particle id is >= max_particles; by-id read-backs "
                                         "need unique ids below max_particles (a permutation of 0..n-1 fills every slot)");
    return SPH_OK;
}

// ---------------------------------------------------------------- slab mode (multi-GPU) ----------
// NCCL is resolved at run time (dlopen) so that the single-GPU library has no link dependency and, in
// a process that already carries torch's bundled NCCL, the very same copy is used.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api(std::string *why) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            auto sym = [&](const char *n) { return dlsym(api.handle, n); };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
            api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
            api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
            api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        }
    }
    const bool ok = api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv &&
                    api.GroupStart && api.GroupEnd && api.GetErrorString;
    if (!ok && why) *why = "NCCL::libnccl.so.2 could not be loaded (slab mode needs NCCL)";
    return ok ? &api : nullptr;
}

#define NCCL_TRY(ctx, api, call)                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(ctx, SPH_ERR_COMM, std::string("NCCL::") + (api)->GetErrorString(r_) + " | " #call); \
    } while (0)

void slab_release(sph_context *c) {
    Slab *s = c->slab;
    if (!s) return;
    if (s->comm) {
        if (NcclApi *api = nccl_api(nullptr)) api->CommDestroy(s->comm);
    }
    if (s->local) {
        // wake neighbours that may be blocked on this rank, then drop the shared mailbox with the last rank
        LocalComm *lc = s->local;
        for (LocalEdge *e : {s->rank + 1 < s->world ? &lc->up[s->rank] : nullptr, s->rank > 0 ? &lc->down[s->rank] : nullptr,
                             s->rank > 0 ? &lc->up[s->rank - 1] : nullptr, s->rank + 1 < s->world ? &lc->down[s->rank + 1] : nullptr})
            if (e) {
                std::lock_guard<std::mutex> g(e->m);
                e->aborted = true;
                e->cv.notify_all();
            }
        if (lc->refs.fetch_sub(1) == 1) {
            for (int r = 0; r < lc->world; ++r)
                for (LocalEdge *e : {&lc->up[r], &lc->down[r]}) {
                    if (e->packed) cudaEventDestroy(e->packed);
                    if (e->copied) cudaEventDestroy(e->copied);
                }
            delete lc;
        }
    }
    for (void *p : {(void *)s->down_pos, (void *)s->down_vel, (void *)s->up_pos, (void *)s->up_vel, (void *)s->d_counters, (void *)s->d_aux})
        if (p) cudaFree(p);
    for (auto &pair : s->ev_busy)
        for (cudaEvent_t e : pair)
            if (e) cudaEventDestroy(e);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (cudaEvent_t e : {s->ev_ranges, s->ev_boundary, s->ev_counts, s->ev_comm})
        if (e) cudaEventDestroy(e);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    delete s;
    c->slab = nullptr;
}

// Layers [z0, z1) of rank `rank`: equal split, the remainder goes to the lowest ranks.
void slab_plan(int rz, int world, int rank, int *z0, int *z1) {
    const int base = rz / world, extra = rz % world;
    *z0 = rank * base + std::min(rank, extra);
    *z1 = *z0 + base + (rank < extra ? 1 : 0);
}

// Ghost exchange + migration in one message pair per face.  Every rank sends, to each neighbour, its
// particles lying in the two layers on either side of the shared face (ghosts for the neighbour) or beyond
// it (migrants); the receiver appends them behind its own particles and the ordinary grid build sorts all.
//
// The exchange is split in three stream-ordered pieces so that the overlapped step can run them on the
// communication stream: pack (one or two index ranges of A), counts (4-byte messages, then read back),
// payload (exact sizes, received straight into the tail of A).
void slab_pack_begin(sph_context *c, cudaStream_t st) {
    Slab &s = *c->slab;
    if (s.local) {
        // the face buffers are about to be overwritten: the neighbours must have copied the previous message out
        for (LocalEdge *e : {s.rank + 1 < s.world ? &s.local->up[s.rank] : nullptr, s.rank > 0 ? &s.local->down[s.rank] : nullptr}) {
            if (!e) continue;
            std::unique_lock<std::mutex> g(e->m);
            e->cv.wait(g, [&] { return e->consumed == e->posted || e->aborted; });
            if (e->posted > 0 && !e->aborted) cudaStreamWaitEvent(st, e->copied, 0);
        }
    }
    cudaMemsetAsync(s.d_counters, 0, 4 * sizeof(int), st);
}

void slab_pack_range(sph_context *c, cudaStream_t st, uint32_t first, uint32_t count) {
    Slab &s = *c->slab;
    const bool has_down = s.rank > 0, has_up = s.rank + 1 < s.world;
    // thresholds relative to the faces of the COMING step (P.next_z0 / next_z1; equal to z0 / z1 unless a face moves)
    launch_slab_pack(c->pos_a + first, c->vel_a + first, (int)count, has_down ? c->P.next_z0 + 2 : -1,
                     has_up ? c->P.next_z1 - 2 : 0x7fffffff, s.down_pos, s.down_vel, s.up_pos, s.up_vel, s.d_counters, s.cap_face,
                     c->P, st);
    c->kernel_launches += 1;
}

// counts[0..1] = particles this rank sends down/up, counts[2..3] = particles it will receive from below/above
int local_exchange_counts(sph_context *c, cudaStream_t st, int counts[4]) {
    Slab &s = *c->slab;
    const bool has_down = s.rank > 0, has_up = s.rank + 1 < s.world;
    CUDA_TRY(c, cudaMemcpyAsync(s.h_pinned, s.d_counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaEventRecord(s.ev_counts, st));
    CUDA_TRY(c, cudaEventSynchronize(s.ev_counts));
    counts[0] = s.h_pinned[0];
    counts[1] = s.h_pinned[1];
    REQUIRE(c, counts[0] <= s.cap_face && counts[1] <= s.cap_face, SPH_ERR_STATE, "slab: face buffer overflow");
    auto post = [&](LocalEdge &e, int count, const float4 *pos, const float4 *vel) -> int {
        std::lock_guard<std::mutex> g(e.m);
        e.count = count;
        std::memcpy(e.aux, s.aux_sent, sizeof(e.aux));
        e.pos = pos;
        e.vel = vel;
        CUDA_TRY(c, cudaEventRecord(e.packed, st));
        e.posted += 1;
        e.cv.notify_all();
        return SPH_OK;
    };
    if (has_down)
        if (int rc = post(s.local->down[s.rank], counts[0], s.down_pos, s.down_vel)) return rc;
    if (has_up)
        if (int rc = post(s.local->up[s.rank], counts[1], s.up_pos, s.up_vel)) return rc;
    auto peek = [&](LocalEdge &e, int *count, int *aux) -> int {
        std::unique_lock<std::mutex> g(e.m);
        e.cv.wait(g, [&] { return e.posted > e.consumed || e.aborted; });
        REQUIRE(c, !e.aborted, SPH_ERR_COMM, "slab (loop-back): a neighbouring rank was destroyed");
        *count = e.count;
        std::memcpy(aux, e.aux, sizeof(e.aux));
        return SPH_OK;
    };
    counts[2] = counts[3] = 0;
    if (has_down)
        if (int rc = peek(s.local->up[s.rank - 1], &counts[2], s.aux_recv[0])) return rc;
    if (has_up)
        if (int rc = peek(s.local->down[s.rank + 1], &counts[3], s.aux_recv[1])) return rc;
    s.aux_valid = true;
    return SPH_OK;
}

int local_exchange_payload(sph_context *c, cudaStream_t st, const int counts[4], uint32_t dst_first) {
    Slab &s = *c->slab;
    const bool has_down = s.rank > 0, has_up = s.rank + 1 < s.world;
    REQUIRE(c, (uint64_t)dst_first + (uint64_t)counts[2] + (uint64_t)counts[3] <= c->cap, SPH_ERR_STATE,
            "slab: local particle capacity exceeded");
    float4 *rp = c->pos_a + dst_first, *rv = c->vel_a + dst_first;
    auto take = [&](LocalEdge &e, float4 *dp, float4 *dv, int count) -> int {
        std::lock_guard<std::mutex> g(e.m);
        CUDA_TRY(c, cudaStreamWaitEvent(st, e.packed, 0));
        if (count) {
            CUDA_TRY(c, cudaMemcpyAsync(dp, e.pos, (size_t)count * sizeof(float4), cudaMemcpyDefault, st));
            CUDA_TRY(c, cudaMemcpyAsync(dv, e.vel, (size_t)count * sizeof(float4), cudaMemcpyDefault, st));
        }
        CUDA_TRY(c, cudaEventRecord(e.copied, st));
        e.consumed += 1;
        e.cv.notify_all();
        return SPH_OK;
    };
    if (has_down)
        if (int rc = take(s.local->up[s.rank - 1], rp, rv, counts[2])) return rc;
    if (has_up)
        if (int rc = take(s.local->down[s.rank + 1], rp + counts[2], rv + counts[2], counts[3])) return rc;
    s.sent_particles += (uint64_t)counts[0] + (uint64_t)counts[1];
    s.exchanges += 1;
    return SPH_OK;
}

int slab_exchange_counts(sph_context *c, cudaStream_t st, int counts[4]) {
    Slab &s = *c->slab;
    if (s.local) return local_exchange_counts(c, st, counts);
    NcclApi *api = nccl_api(&c->err);
    if (!api) return SPH_ERR_COMM;
    const bool has_down = s.rank > 0, has_up = s.rank + 1 < s.world;
    std::memcpy(s.h_pinned + 8, s.aux_sent, 4 * sizeof(int));  // the record that travels with the counts
    CUDA_TRY(c, cudaMemcpyAsync(s.d_aux, s.h_pinned + 8, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(c, api, api->GroupStart());
    if (has_down) {
        NCCL_TRY(c, api, api->Send(s.d_counters + 0, 1, ncclInt32, s.rank - 1, s.comm, st));
        NCCL_TRY(c, api, api->Send(s.d_aux, 4, ncclInt32, s.rank - 1, s.comm, st));
        NCCL_TRY(c, api, api->Recv(s.d_counters + 2, 1, ncclInt32, s.rank - 1, s.comm, st));
        NCCL_TRY(c, api, api->Recv(s.d_aux + 4, 4, ncclInt32, s.rank - 1, s.comm, st));
    }
    if (has_up) {
        NCCL_TRY(c, api, api->Send(s.d_counters + 1, 1, ncclInt32, s.rank + 1, s.comm, st));
        NCCL_TRY(c, api, api->Send(s.d_aux, 4, ncclInt32, s.rank + 1, s.comm, st));
        NCCL_TRY(c, api, api->Recv(s.d_counters + 3, 1, ncclInt32, s.rank + 1, s.comm, st));
        NCCL_TRY(c, api, api->Recv(s.d_aux + 8, 4, ncclInt32, s.rank + 1, s.comm, st));
    }
    NCCL_TRY(c, api, api->GroupEnd());
    CUDA_TRY(c, cudaMemcpyAsync(s.h_pinned, s.d_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_pinned + 12, s.d_aux + 4, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaEventRecord(s.ev_counts, st));
    CUDA_TRY(c, cudaEventSynchronize(s.ev_counts));  // only this stream's work is waited for
    std::memcpy(s.aux_recv[0], s.h_pinned + 12, 4 * sizeof(int));
    std::memcpy(s.aux_recv[1], s.h_pinned + 16, 4 * sizeof(int));
    s.aux_valid = true;
    counts[0] = s.h_pinned[0];
    counts[1] = s.h_pinned[1];
    counts[2] = has_down ? s.h_pinned[2] : 0;
    counts[3] = has_up ? s.h_pinned[3] : 0;
    REQUIRE(c, counts[0] <= s.cap_face && counts[1] <= s.cap_face, SPH_ERR_STATE, "slab: face buffer overflow");
    return SPH_OK;
}

int slab_exchange_payload(sph_context *c, cudaStream_t st, const int counts[4], uint32_t dst_first) {
    Slab &s = *c->slab;
    if (s.local) return local_exchange_payload(c, st, counts, dst_first);
    NcclApi *api = nccl_api(&c->err);
    if (!api) return SPH_ERR_COMM;
    const bool has_down = s.rank > 0, has_up = s.rank + 1 < s.world;
    REQUIRE(c, (uint64_t)dst_first + (uint64_t)counts[2] + (uint64_t)counts[3] <= c->cap, SPH_ERR_STATE,
            "slab: local particle capacity exceeded");
    float4 *rp = c->pos_a + dst_first, *rv = c->vel_a + dst_first;
    NCCL_TRY(c, api, api->GroupStart());
    if (has_down) {
        if (counts[0]) {
            NCCL_TRY(c, api, api->Send(s.down_pos, (size_t)counts[0] * 4, ncclFloat32, s.rank - 1, s.comm, st));
            NCCL_TRY(c, api, api->Send(s.down_vel, (size_t)counts[0] * 4, ncclFloat32, s.rank - 1, s.comm, st));
        }
        if (counts[2]) {
            NCCL_TRY(c, api, api->Recv(rp, (size_t)counts[2] * 4, ncclFloat32, s.rank - 1, s.comm, st));
            NCCL_TRY(c, api, api->Recv(rv, (size_t)counts[2] * 4, ncclFloat32, s.rank - 1, s.comm, st));
        }
    }
    if (has_up) {
        if (counts[1]) {
            NCCL_TRY(c, api, api->Send(s.up_pos, (size_t)counts[1] * 4, ncclFloat32, s.rank + 1, s.comm, st));
            NCCL_TRY(c, api, api->Send(s.up_vel, (size_t)counts[1] * 4, ncclFloat32, s.rank + 1, s.comm, st));
        }
        if (counts[3]) {
            NCCL_TRY(c, api, api->Recv(rp + counts[2], (size_t)counts[3] * 4, ncclFloat32, s.rank + 1, s.comm, st));
            NCCL_TRY(c, api, api->Recv(rv + counts[2], (size_t)counts[3] * 4, ncclFloat32, s.rank + 1, s.comm, st));
        }
    }
    NCCL_TRY(c, api, api->GroupEnd());
    s.sent_particles += (uint64_t)counts[0] + (uint64_t)counts[1];
    s.exchanges += 1;
    return SPH_OK;
}

// Faces of the coming exchange, from the records both neighbours swapped with the previous one (see Slab).
void slab_plan_faces(sph_context *c) {
    Slab &s = *c->slab;
    int nz0 = s.z0, nz1 = s.z1;
    if (s.opt_rebalance && s.aux_valid && s.world > 1) {
        if (s.rank > 0)
            nz0 += face_shift(s.aux_recv[0], s.aux_sent, s.z0 - s.z0_init, s.shift_max, s.opt_rebalance, s.exchanges, s.rank - 1);
        if (s.rank + 1 < s.world)
            nz1 += face_shift(s.aux_sent, s.aux_recv[1], s.z1 - s.z1_init, s.shift_max, s.opt_rebalance, s.exchanges, s.rank);
    }
    c->P.next_z0 = nz0;
    c->P.next_z1 = nz1;
}

// The record that travels with the coming exchange, and the switch to the new faces once it is under way.
void slab_fill_aux(sph_context *c) {
    Slab &s = *c->slab;
    const long long room = (long long)c->cap - 2LL * s.cap_face;
    s.aux_sent[0] = s.load_us;
    s.aux_sent[1] = c->P.next_z0;
    s.aux_sent[2] = c->P.next_z1;
    s.aux_sent[3] = (long long)s.n_own * 10 < room * 9 ? 1 : 0;
}
// The part of the local grid the owned layers [z0, z1) put to use.  Particles live in the owned layers and two ghost
// layers per face; sort keys are confined to one layer more on either side (a particle that is further out is an error
// the exchange reports, a non-finite one must merely land in a valid bin), the scan covers yet another layer, so every
// candidate window of every local particle reads scanned bins.  The spare layers that give the faces room to move
// (shift_max) cost nothing per step this way.  Densities: owned layers + the inner ghost layer of each face.
void slab_set_layer_windows(sph_context *c) {
    Slab &s = *c->slab;
    Params &P = c->P;
    const int l0 = s.z0 - s.z_base, l3 = s.z1 - s.z_base;
    const long long plane = (long long)P.rx * P.xb * P.ry;  // sort-key entries per z-layer
    P.kz_lo = std::max(l0 - 3, 0);
    P.kz_hi = std::min(l3 + 3, P.rz);
    P.dens_key_lo = (int)(std::max(l0 - 1, 0) * plane);
    P.dens_key_hi = (int)std::min<long long>(std::min(l3 + 1, P.rz) * plane, 0x7fffffffLL);
    const long long lo = std::max(P.kz_lo - 1, 0) * plane, hi = std::min(P.kz_hi + 1, P.rz) * plane + 1;  // bins [lo, hi)
    c->g.scan_tile0 = (int)(lo / kScanTile);
    c->g.scan_tiles = std::min(c->g.n_tiles, (int)((hi + kScanTile - 1) / kScanTile)) - c->g.scan_tile0;
}
void slab_adopt_faces(sph_context *c) {
    Slab &s = *c->slab;
    if (c->P.next_z0 != s.z0) s.face_moves += 1;
    if (c->P.next_z1 != s.z1) s.face_moves += 1;
    s.z0 = c->P.own_z0 = c->P.next_z0;
    s.z1 = c->P.own_z1 = c->P.next_z1;
    slab_set_layer_windows(c);
}

// Blocking exchange on the compute stream: pack the whole carried view, then counts, then payload.
int slab_exchange(sph_context *c) {
    Slab &s = *c->slab;
    int counts[4];
    slab_plan_faces(c);
    slab_fill_aux(c);
    slab_pack_begin(c, c->stream);
    slab_pack_range(c, c->stream, c->in_off, s.n_own);
    int rc = slab_exchange_counts(c, c->stream, counts);
    if (rc) return rc;
    rc = slab_exchange_payload(c, c->stream, counts, c->in_off + s.n_own);
    if (rc) return rc;
    slab_adopt_faces(c);
    c->n = s.n_own + (uint32_t)counts[2] + (uint32_t)counts[3];
    s.have_ghosts = true;
    return SPH_OK;
}

// Read back the first particle index of up to 4 local z-layers (cell_start at layer boundaries).
int slab_layer_starts(sph_context *c, const int layers[4], int out[4]) {
    Slab &s = *c->slab;
    (void)out;
    const size_t rxy = (size_t)c->P.rx * c->P.xb * c->P.ry;  // sort-key entries per z-layer
    launch_gather4(c->g.cell_start, (size_t)layers[0] * rxy, (size_t)layers[1] * rxy, (size_t)layers[2] * rxy,
                   (size_t)layers[3] * rxy, s.h_pinned + 4, c->stream);  // zero-copy store into the pinned words
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaEventRecord(s.ev_ranges, c->stream));
    return SPH_OK;
}

// busy time of a complete step (microseconds) from its event pair; the load figure the neighbours compare
void slab_read_load(Slab &s, int parity) {
    if (!s.busy_recorded[parity]) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.ev_busy[parity][0], s.ev_busy[parity][1]) == cudaSuccess) s.load_us = (int)(ms * 1000.0f) + 1;
    s.busy_recorded[parity] = 0;
}

// Sequential slab step (exchange, then the ordinary step) — reference behaviour for the overlapped variant.
int slab_step_sequential(sph_context *c, int n_steps, double *ms) {
    Slab &s = *c->slab;
    PhaseTimer t(c, ms);
    for (int k = 0; k < n_steps; ++k) {
        if (!s.have_ghosts) {
            int rc = slab_exchange(c);
            if (rc) return rc;
        }
        c->P.next_z0 = s.z0;  // no face moves inside the step: the next exchange plans them
        c->P.next_z1 = s.z1;
        CUDA_TRY(c, cudaEventRecord(s.ev_busy[0][0], c->stream));
        enqueue_step(c);
        CUDA_TRY(c, cudaEventRecord(s.ev_busy[0][1], c->stream));
        s.busy_recorded[0] = 1;
        s.have_ghosts = false;
        // the owned particles are the contiguous run of the owned layers in the canonical order
        const int layers[4] = {s.z0 - s.z_base, s.z0 - s.z_base, s.z1 - s.z_base, s.z1 - s.z_base};
        int rc = slab_layer_starts(c, layers, nullptr);
        if (rc) return rc;
        CUDA_TRY(c, cudaEventSynchronize(s.ev_ranges));
        slab_read_load(s, 0);
        c->in_off = (uint32_t)s.h_pinned[4];
        s.n_own = (uint32_t)(s.h_pinned[6] - s.h_pinned[4]);
    }
    c->steps += (uint64_t)n_steps;
    c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
    int rc = check_launch(c, "slab step");
    return rc ? rc : t.finish();
}

// Overlapped slab step.  After the density pass the forces + integration of the boundary layers (the owned layers next
// to each face, down to four layers inside the face this rank will own after the exchange: two that become the
// neighbour's ghosts plus two of slack for particles moving towards the face) run on the high-priority communication
// stream, followed there by the pack and the NCCL exchange for the NEXT step; the interior runs concurrently on the
// compute stream.  Stream priority matters: without it the interior kernel's queued blocks keep every SM slot and
// the pack only starts when the interior has drained (measured: profiles/r1_timeline_slab_2gpu.txt).  Ghosts are
// neither force-evaluated nor integrated, so the received particles can land directly behind the owned run of A
// while the interior is still being written.
int slab_step_overlapped(sph_context *c, int n_steps, double *ms) {
    Slab &s = *c->slab;
    constexpr int kBoundaryLayers = 4;
    PhaseTimer t(c, ms);
    for (int k = 0; k < n_steps; ++k) {
        if (!s.have_ghosts) {  // first step after an upload: nothing to overlap with yet
            int rc = slab_exchange(c);
            if (rc) return rc;
        }
        const int parity = (int)(c->steps + (uint64_t)k) & 1;
        enqueue_grid(c);
        slab_plan_faces(c);  // the faces after this step's exchange: the kernels below see them as P.next_z0 / next_z1
        const int l0 = s.z0 - s.z_base, l3 = s.z1 - s.z_base;
        const int l1 = std::min(std::max(s.z0, c->P.next_z0) - s.z_base + kBoundaryLayers, l3);
        const int l2 = std::max(std::min(s.z1, c->P.next_z1) - s.z_base - kBoundaryLayers, l1);
        const int layers[4] = {l0, l1, l2, l3};
        int rc = slab_layer_starts(c, layers, nullptr);
        if (rc) return rc;
        CUDA_TRY(c, cudaEventRecord(s.ev_busy[parity][0], c->stream));
        enqueue_density(c);  // all local particles: the first ghost layer needs its density too
        CUDA_TRY(c, cudaEventRecord(s.ev_boundary, c->stream));  // "density done": the boundary work may start
        CUDA_TRY(c, cudaEventSynchronize(s.ev_ranges));  // returns while the density pass is still running
        slab_read_load(s, parity ^ 1);                   // the previous step is complete: its busy time is final
        const int L0 = s.h_pinned[4], L1 = s.h_pinned[5], L2 = s.h_pinned[6], L3 = s.h_pinned[7];
        auto forces_integrate = [&](int i0, int i1, cudaStream_t st) {  // fused forces + walls + integration on [i0, i1)
            launch_forces_mask(c->nb, c->dp, c->nb_count, c->g.key_s, c->g.cell_start, c->acc, i0, i1, c->P, st,
                               c->pos_s, c->pos_a, c->vel_a, s.d_counters + 4);
            c->kernel_launches += 2;
        };
        // ---- boundary layers, pack and exchange for the next step: high-priority stream
        CUDA_TRY(c, cudaStreamWaitEvent(s.comm_stream, s.ev_boundary, 0));
        forces_integrate(L0, L1, s.comm_stream);
        forces_integrate(L2, L3, s.comm_stream);
        // ---- interior: compute stream, concurrently (its blocks fill whatever the boundary work leaves free)
        forces_integrate(L1, L2, c->stream);
        CUDA_TRY(c, cudaEventRecord(s.ev_busy[parity][1], c->stream));
        s.busy_recorded[parity] = 1;
        slab_fill_aux(c);
        slab_pack_begin(c, s.comm_stream);
        slab_pack_range(c, s.comm_stream, (uint32_t)L0, (uint32_t)(L1 - L0));
        slab_pack_range(c, s.comm_stream, (uint32_t)L2, (uint32_t)(L3 - L2));
        int counts[4];
        rc = slab_exchange_counts(c, s.comm_stream, counts);
        if (rc) return rc;
        rc = slab_exchange_payload(c, s.comm_stream, counts, (uint32_t)L3);
        if (rc) return rc;
        CUDA_TRY(c, cudaEventRecord(s.ev_comm, s.comm_stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, s.ev_comm, 0));  // the next grid build needs the received tail
        slab_adopt_faces(c);
        c->in_off = (uint32_t)L0;  // carried forward: the run this rank owned during the step (a layer it has just
        s.n_own = (uint32_t)(L3 - L0);  // given away stays as its own, current, ghost copy; the next sort reclassifies)
        c->n = s.n_own + (uint32_t)counts[2] + (uint32_t)counts[3];
        s.have_ghosts = true;
    }
    c->steps += (uint64_t)n_steps;
    c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
    int rc = check_launch(c, "slab step");
    return rc ? rc : t.finish();
}

int slab_step(sph_context *c, int n_steps, double *ms) {
    if (c->slab->opt_overlap && use_mask_passes(c) && (c->slab->z1 - c->slab->z0) >= 8)
        return slab_step_overlapped(c, n_steps, ms);
    return slab_step_sequential(c, n_steps, ms);
}

}  // namespace

extern "C" {

int sph_abi_version(void) { return SPH_ABI_VERSION; }

int sph_device_count(int *count) {
    REQUIRE(nullptr, count, SPH_ERR_ARGUMENT, "sph_device_count: count is NULL");
    *count = 0;
    CUDA_TRY(nullptr, cudaGetDeviceCount(count));
    return SPH_OK;
}

int sph_device_name(int device, char *buf, size_t len) {
    REQUIRE(nullptr, buf && len > 0, SPH_ERR_ARGUMENT, "sph_device_name: empty buffer");
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    std::snprintf(buf, len, "%s (CUDA sm_%d%d, %d SMs, %.0f GB)", prop.name, prop.major, prop.minor,
                  prop.multiProcessorCount, (double)prop.totalGlobalMem / 1e9);
    return SPH_OK;
}

int sph_config_init(sph_config *cfg, float bx, float by, float bz, uint32_t max_particles) {
    REQUIRE(nullptr, cfg, SPH_ERR_ARGUMENT, "sph_config_init: cfg is NULL");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->box[0] = bx;
    cfg->box[1] = by;
    cfg->box[2] = bz;
    cfg->h = 0.0457f;             // include/CParticle.h:80
    cfg->viscosity = 3.5f;        // :81
    cfg->mass = 0.02f;            // :82
    cfg->gas_stiffness = 3.0f;    // :83
    cfg->rest_density = 998.29f;  // :84
    cfg->dt = 0.01f;              // src/CBaseParticleSimulator.cpp:7
    for (int a = 0; a < 3; ++a) cfg->grid_res[a] = (int)std::ceil(cfg->box[a] / cfg->h);  // :27-31 (float division)
    cfg->gravity[1] = -9.80665f;  // include/CBaseParticleSimulator.h:19
    cfg->wall_k = 10000.0;        // include/CCollisionGeometry.h:20
    cfg->wall_damping = -0.9;     // :21
    cfg->wall_skin = 0.01;
    cfg->wall_count = 6;
    const float mn[3] = {-(bx / 2.0f), -(by / 2.0f), -(bz / 2.0f)}, mx[3] = {bx / 2.0f, by / 2.0f, bz / 2.0f};
    for (int a = 0; a < 3; ++a) {  // include/CCollisionGeometry.h:79-120
        cfg->walls[a].normal[a] = -1.0f;
        cfg->walls[a].position[a] = mn[a];
        cfg->walls[a + 3].normal[a] = 1.0f;
        cfg->walls[a + 3].position[a] = mx[a];
    }
    cfg->max_particles = max_particles;
    cfg->device = 0;
    cfg->rank = 0;
    cfg->world = 1;
    return SPH_OK;
}

int sph_destroy(sph_context *c) {
    if (!c) return SPH_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graph(c);
    void *ptrs[] = {c->pos_a, c->vel_a, c->pos_s, c->vel_s, c->dp, c->acc, c->nb_count, c->g.key_a, c->g.off_a,
                    c->g.bucket_src, c->g.bucket_id, c->g.key_s, c->g.count, c->g.cell_start, c->g.scan_status,
                    c->d_stage, c->d_tmp_i32, c->d_tmp_f32, c->d_stats, c->d_id_error, c->d_hist, c->nb.xs, c->nb.ys, c->nb.zs, c->nb.fdat, c->nb.mask, c->nb.words, c->nb.ovf};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->d_flush) cudaFree(c->d_flush);
    if (c->d_mesh_planes) cudaFree(c->d_mesh_planes);
    for (int k = 0; k < 2; ++k) {
        if (c->snap_pos[k]) cudaFree(c->snap_pos[k]);
        if (c->snap_vel[k]) cudaFree(c->snap_vel[k]);
    }
    if (c->d_emit_pos) cudaFree(c->d_emit_pos);
    if (c->d_emit_vel) cudaFree(c->d_emit_vel);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    slab_release(c);
    for (cudaEvent_t e : c->step_events) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->pinned_ptr) cudaHostUnregister(c->pinned_ptr);
    delete c;
    return SPH_OK;
}

int sph_create(const sph_config *cfg, sph_context **out) {
    REQUIRE(nullptr, cfg && out, SPH_ERR_ARGUMENT, "sph_create: NULL argument");
    *out = nullptr;
    REQUIRE(nullptr, cfg->grid_res[0] > 0 && cfg->grid_res[1] > 0 && cfg->grid_res[2] > 0, SPH_ERR_ARGUMENT,
            "sph_create: grid resolution must be positive");
    REQUIRE(nullptr, (double)cfg->grid_res[0] * cfg->grid_res[1] * cfg->grid_res[2] < 2.0e9, SPH_ERR_ARGUMENT,
            "sph_create: too many cells for 32-bit keys");
    REQUIRE(nullptr, cfg->max_particles > 0 && cfg->max_particles < 0x7fffff00u, SPH_ERR_ARGUMENT,
            "sph_create: max_particles out of range");
    REQUIRE(nullptr, cfg->wall_count >= 0 && cfg->wall_count <= 6, SPH_ERR_ARGUMENT, "sph_create: wall_count > 6");
    REQUIRE(nullptr, cfg->world <= 1, SPH_ERR_ARGUMENT, "sph_create: slab mode is created through sph_slab_create");
    int count = 0;
    CUDA_TRY(nullptr, cudaGetDeviceCount(&count));
    REQUIRE(nullptr, cfg->device >= 0 && cfg->device < count, SPH_ERR_CUDA, "sph_create: no such CUDA device");
    CUDA_TRY(nullptr, cudaSetDevice(cfg->device));

    sph_context *c = new sph_context();
    c->cfg = *cfg;
    c->device = cfg->device;
    c->P = make_params(*cfg);
    // x bins per cell of the sort key: 4 by default (the density pass then scans 9 of 12 bins per row); 1 for grids
    // too narrow for the bitmask passes; SPH_XBINS overrides for experiments (1, 2, 4 or 8)
    int xb = c->P.rx >= 4 ? 4 : 1;
    if (const char *e = std::getenv("SPH_XBINS")) {
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) xb = v;
    }
    while (xb > 1 && (double)c->P.n_cells * xb >= 2.0e9) xb >>= 1;
    c->P.xb = xb;
    c->cap = cfg->max_particles;
    const size_t cap = c->cap;
    const size_t items = (size_t)c->P.n_cells * (size_t)xb + 1;
    c->g.n_scan_items = (int)items;
    c->g.n_tiles = (int)((items + kScanTile - 1) / kScanTile);
    c->g.scan_tile0 = 0;
    c->g.scan_tiles = c->g.n_tiles;
    c->cells_padded = (size_t)c->g.n_tiles * kScanTile;

#define CTX_TRY(call)                       \
    do {                                    \
        cudaError_t e_ = (call);            \
        if (e_ != cudaSuccess) {            \
            std::string m = std::string("CUDA::") + cudaGetErrorName(e_) + " | " #call; \
            sph_destroy(c);                 \
            return fail(nullptr, SPH_ERR_CUDA, m); \
        }                                   \
    } while (0)

    CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CTX_TRY(cudaEventCreate(&c->ev0));
    CTX_TRY(cudaEventCreate(&c->ev1));
    CTX_TRY(dalloc(&c->pos_a, cap));
    CTX_TRY(dalloc(&c->vel_a, cap));
    CTX_TRY(dalloc(&c->pos_s, cap));
    CTX_TRY(dalloc(&c->vel_s, cap));
    CTX_TRY(dalloc(&c->dp, cap));
    CTX_TRY(dalloc(&c->acc, cap));
    CTX_TRY(dalloc(&c->nb_count, cap));
    CTX_TRY(dalloc(&c->g.key_a, cap));
    CTX_TRY(dalloc(&c->g.off_a, cap));
    CTX_TRY(dalloc(&c->g.bucket_src, cap));
    CTX_TRY(dalloc(&c->g.bucket_id, cap));
    CTX_TRY(dalloc(&c->g.key_s, cap));
    CTX_TRY(dalloc(&c->g.count, c->cells_padded));
    CTX_TRY(dalloc(&c->g.cell_start, c->cells_padded));
    CTX_TRY(dalloc(&c->g.scan_status, (size_t)c->g.n_tiles + 1));
    CTX_TRY(dalloc(&c->nb.xs, cap + kSoaPad));
    CTX_TRY(dalloc(&c->nb.ys, cap + kSoaPad));
    CTX_TRY(dalloc(&c->nb.zs, cap + kSoaPad));
    CTX_TRY(dalloc(&c->nb.fdat, 2 * cap));
    CTX_TRY(dalloc(&c->nb.mask, ((cap + 31) / 32) * (size_t)kMaskWords * 32));
    CTX_TRY(dalloc(&c->nb.words, cap));
    CTX_TRY(dalloc(&c->nb.ovf, cap + 1));
    CTX_TRY(dalloc(&c->d_tmp_i32, cap));
    CTX_TRY(dalloc(&c->d_tmp_f32, cap * 5));
    CTX_TRY(dalloc(&c->d_stats, (size_t)8));
    CTX_TRY(dalloc(&c->d_id_error, (size_t)1));
    CTX_TRY(cudaMemsetAsync(c->d_id_error, 0, sizeof(int), c->stream));
    CTX_TRY(dalloc(&c->d_hist, (size_t)kHistBins));
    c->stage_cap = std::min<size_t>(cap, (size_t)4 << 20);  // <= 4 Mi records (320 MiB) of AoS staging
    CTX_TRY(dalloc(&c->d_stage, c->stage_cap));
    CTX_TRY(cudaMemsetAsync(c->g.count, 0, c->cells_padded * sizeof(int), c->stream));
    CTX_TRY(cudaMemsetAsync(c->g.scan_status, 0, ((size_t)c->g.n_tiles + 1) * sizeof(unsigned long long), c->stream));
    CTX_TRY(cudaMemsetAsync(c->nb.ovf, 0, sizeof(int), c->stream));
    CTX_TRY(cudaMemsetAsync(c->g.cell_start, 0, c->cells_padded * sizeof(int), c->stream));
    CTX_TRY(cudaMemsetAsync(c->acc, 0, cap * sizeof(float4), c->stream));
    CTX_TRY(cudaMemsetAsync(c->dp, 0, cap * sizeof(float4), c->stream));
    CTX_TRY(cudaMemsetAsync(c->nb_count, 0, cap * sizeof(int), c->stream));
    CTX_TRY(cudaStreamSynchronize(c->stream));
#undef CTX_TRY
    *out = c;
    return SPH_OK;
}

const char *sph_last_error(const sph_context *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

// ---------------------------------------------------------------- state
// d_stage is shared by uploads and read-backs: an asynchronous copy still in flight has to drain first
static int drain_async_download(sph_context *c) {
    if (c->copy_pending) {
        CUDA_TRY(c, cudaEventSynchronize(c->ev_copy));
        c->copy_pending = false;
    }
    return SPH_OK;
}

static int upload_range(sph_context *c, const sph_particle *aos, uint32_t first, uint32_t count) {
    if (int rc = drain_async_download(c)) return rc;
    for (uint32_t done = 0; done < count;) {
        const uint32_t chunk = (uint32_t)std::min<size_t>(count - done, c->stage_cap);
        CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, aos + done, (size_t)chunk * sizeof(sph_particle), cudaMemcpyHostToDevice,
                                    c->stream));
        launch_aos_to_soa(c->d_stage, c->pos_a + first + done, c->vel_a + first + done, (int)chunk,
                          c->slab ? 0xffffffffu : c->cap, c->d_id_error, c->stream);  // slab contexts carry global ids
        c->kernel_launches += 1;
        done += chunk;
        if (done < count) CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // staging buffer is reused
    }
    return check_launch(c, "aos_to_soa");
}

int sph_upload_particles(sph_context *c, const sph_particle *aos, uint32_t n) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, n <= c->cap, SPH_ERR_ARGUMENT, "sph_upload_particles: n exceeds max_particles");
    REQUIRE(c, aos || n == 0, SPH_ERR_ARGUMENT, "sph_upload_particles: NULL particles");
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->n = n;
    c->in_off = 0;
    if (c->slab) {
        c->slab->n_own = n;
        c->slab->have_ghosts = false;
    }
    c->s_valid = c->grid_valid = c->a_aligned = c->density_valid = c->forces_valid = false;
    CUDA_TRY(c, cudaMemsetAsync(c->d_id_error, 0, sizeof(int), c->stream));  // the whole state is replaced
    return upload_range(c, aos, 0, n);
}

int sph_append_particles(sph_context *c, const sph_particle *aos, uint32_t n_new) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_append_particles: not available in slab mode");
    REQUIRE(c, (uint64_t)c->n + n_new <= c->cap, SPH_ERR_ARGUMENT, "sph_append_particles: exceeds max_particles");
    REQUIRE(c, aos || n_new == 0, SPH_ERR_ARGUMENT, "sph_append_particles: NULL particles");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const uint32_t first = c->n;
    c->n += n_new;
    // A stays authoritative; the snapshot S and everything derived from it no longer covers all particles
    c->s_valid = c->grid_valid = c->a_aligned = c->density_valid = c->forces_valid = false;
    return upload_range(c, aos, first, n_new);
}

// Snapshot / restore of the authoritative state on the device (two slots): positions, velocities and their current
// array order, so that a run can be replayed EXACTLY — same particle order in memory, hence the same memory access
// pattern and timing — e.g. to time the per-kernel split on the very steps whose total was timed before.
int sph_state_save(sph_context *c, int slot) {
    REQUIRE(c, c && (slot == 0 || slot == 1), SPH_ERR_ARGUMENT, "sph_state_save: slot must be 0 or 1");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_state_save: not available in slab mode");
    REQUIRE(c, c->a_aligned || !c->s_valid, SPH_ERR_STATE, "sph_state_save: call between steps (after integrate)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->snap_pos[slot]) {
        CUDA_TRY(c, dalloc(&c->snap_pos[slot], (size_t)c->cap));
        CUDA_TRY(c, dalloc(&c->snap_vel[slot], (size_t)c->cap));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->snap_pos[slot], c->pos_a + c->in_off, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->snap_vel[slot], c->vel_a + c->in_off, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    c->snap_n[slot] = c->n;
    return SPH_OK;
}

int sph_state_restore(sph_context *c, int slot) {
    REQUIRE(c, c && (slot == 0 || slot == 1), SPH_ERR_ARGUMENT, "sph_state_restore: slot must be 0 or 1");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_state_restore: not available in slab mode");
    REQUIRE(c, c->snap_pos[slot], SPH_ERR_STATE, "sph_state_restore: nothing saved in this slot");
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->n = c->snap_n[slot];
    c->in_off = 0;
    CUDA_TRY(c, cudaMemcpyAsync(c->pos_a, c->snap_pos[slot], (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->vel_a, c->snap_vel[slot], (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    c->s_valid = c->grid_valid = c->a_aligned = c->density_valid = c->forces_valid = false;
    return SPH_OK;
}

int sph_set_emitter(sph_context *c, const sph_particle *templates, uint32_t n_templates, uint32_t group, uint32_t max_count) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_set_emitter: not available in slab mode");
    REQUIRE(c, n_templates == 0 || (templates && group > 0 && n_templates % group == 0), SPH_ERR_ARGUMENT,
            "sph_set_emitter: templates must come in whole groups");
    REQUIRE(c, max_count <= c->cap, SPH_ERR_ARGUMENT, "sph_set_emitter: max_count exceeds max_particles");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->d_emit_pos) cudaFree(c->d_emit_pos), c->d_emit_pos = nullptr;
    if (c->d_emit_vel) cudaFree(c->d_emit_vel), c->d_emit_vel = nullptr;
    c->emit_templates = c->emit_group = c->emit_max = 0;
    if (n_templates == 0) return SPH_OK;
    std::vector<float4> pos(n_templates), vel(n_templates);
    for (uint32_t t = 0; t < n_templates; ++t) {
        pos[t] = make_float4(templates[t].position[0], templates[t].position[1], templates[t].position[2], 0.f);
        vel[t] = make_float4(templates[t].velocity[0], templates[t].velocity[1], templates[t].velocity[2], 0.f);
    }
    CUDA_TRY(c, dalloc(&c->d_emit_pos, (size_t)n_templates));
    CUDA_TRY(c, dalloc(&c->d_emit_vel, (size_t)n_templates));
    CUDA_TRY(c, cudaMemcpy(c->d_emit_pos, pos.data(), n_templates * sizeof(float4), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_emit_vel, vel.data(), n_templates * sizeof(float4), cudaMemcpyHostToDevice));
    c->emit_templates = n_templates;
    c->emit_group = group;
    c->emit_max = max_count;
    return SPH_OK;
}

// One emission (the generateParticles() call at the top of step(), src/CBaseParticleSimulator.cpp:120,187-210): group g of
// the templates is appended while count < max_count - group, ids continue the running count.  Nothing crosses PCIe.
static uint32_t enqueue_emit(sph_context *c) {
    if (!c->emit_templates) return 0;
    uint32_t groups = 0;
    const uint32_t total = c->emit_templates / c->emit_group;
    while (groups < total && c->emit_max >= c->emit_group && c->n + groups * c->emit_group < c->emit_max - c->emit_group &&
           c->n + (groups + 1) * c->emit_group <= c->cap)
        ++groups;
    const uint32_t n_new = groups * c->emit_group;
    if (n_new == 0) return 0;
    launch_emit(c->d_emit_pos, c->d_emit_vel, (int)n_new, c->pos_a + c->in_off + c->n, c->vel_a + c->in_off + c->n, c->n, c->stream);
    c->kernel_launches += 1;
    c->n += n_new;
    c->s_valid = c->grid_valid = c->a_aligned = c->density_valid = c->forces_valid = false;
    return n_new;
}

int sph_emit(sph_context *c, uint32_t *n_emitted) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const uint32_t k = enqueue_emit(c);
    if (n_emitted) *n_emitted = k;
    return check_launch(c, "emit");
}

int sph_download_particles(sph_context *c, sph_particle *aos, uint32_t capacity, uint32_t *n_out) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, aos && capacity >= c->n, SPH_ERR_ARGUMENT, "sph_download_particles: buffer too small");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_particles: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (int rc = drain_async_download(c)) return rc;
    if (int rc = id_precondition(c, "sph_download_particles")) return rc;
    const bool aux = aux_aligned(c);
    for (uint32_t base = 0; base < c->n;) {
        const uint32_t chunk = (uint32_t)std::min<size_t>(c->n - base, c->stage_cap);
        launch_soa_to_aos(view_pos(c), view_vel(c), aux_acc(c), aux_dp(c),
                          (aux && c->grid_valid) ? c->g.key_s : nullptr, c->d_stage, (int)base, (int)chunk, (int)c->n,
                          c->P, c->stream);
        c->kernel_launches += 1;
        CUDA_TRY(c, cudaMemcpyAsync(aos + base, c->d_stage, (size_t)chunk * sizeof(sph_particle), cudaMemcpyDeviceToHost,
                                    c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        base += chunk;
    }
    if (n_out) *n_out = c->n;
    return check_launch(c, "soa_to_aos");
}

int sph_download_particles_async(sph_context *c, sph_particle *aos, uint32_t capacity) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, aos && capacity >= c->n, SPH_ERR_ARGUMENT, "sph_download_particles_async: buffer too small");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_particles_async: not available in slab mode (use sph_download_owned)");
    REQUIRE(c, c->n <= c->stage_cap, SPH_ERR_ARGUMENT, "sph_download_particles_async: more particles than the staging buffer holds; use sph_download_particles");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->copy_stream) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    }
    // the previous copy must have left the staging buffer before the snapshot kernel overwrites it (device-side wait)
    if (c->copy_pending) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    c->copy_n = c->n;
    if (c->n > 0) {
        const bool aux = aux_aligned(c);
        launch_soa_to_aos(view_pos(c), view_vel(c), aux_acc(c), aux_dp(c),
                          (aux && c->grid_valid) ? c->g.key_s : nullptr, c->d_stage, 0, (int)c->n, (int)c->n, c->P, c->stream);
        c->kernel_launches += 1;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_snap, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    if (c->n > 0)
        CUDA_TRY(c, cudaMemcpyAsync(aos, c->d_stage, (size_t)c->n * sizeof(sph_particle), cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_TRY(c, cudaEventRecord(c->ev_copy, c->copy_stream));
    c->copy_pending = true;
    return check_launch(c, "soa_to_aos (async)");
}

int sph_download_wait(sph_context *c, uint32_t *n_out) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (int rc = drain_async_download(c)) return rc;
    if (n_out) *n_out = c->copy_n;
    return SPH_OK;
}

int sph_particle_count(const sph_context *c, uint32_t *n_out) {
    if (!c || !n_out) return SPH_ERR_ARGUMENT;
    *n_out = c->slab ? c->slab->n_own : c->n;
    return SPH_OK;
}

int sph_set_gravity(sph_context *c, float gx, float gy, float gz) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    c->cfg.gravity[0] = c->P.gx = gx;
    c->cfg.gravity[1] = c->P.gy = gy;
    c->cfg.gravity[2] = c->P.gz = gz;
    drop_graph(c);  // kernel parameters are baked into the graph
    return SPH_OK;
}

int sph_set_collision_faces(sph_context *c, const sph_face *faces, uint32_t n_faces) {
    REQUIRE(c, c && (faces || n_faces == 0), SPH_ERR_ARGUMENT, "sph_set_collision_faces: NULL argument");
    REQUIRE(c, n_faces <= (1u << 20), SPH_ERR_ARGUMENT, "sph_set_collision_faces: more than 2^20 faces");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drop_graph(c);
    if (c->d_mesh_planes) {
        cudaFree(c->d_mesh_planes);
        c->d_mesh_planes = nullptr;
    }
    c->P.mesh_planes = nullptr;
    c->P.mesh_plane_count = 0;
    c->P.mesh_k_f = (float)5000.0;  // src/CCollisionGeometry.cpp:109
    if (n_faces == 0) return SPH_OK;
    // The per-face part of inverseBounce is hoisted out of the particle loop: inverseNormal = normal * (-1), then
    // QVector3D::normalize() as Qt 5 defines it (fp64 squared length, early-out when it is fuzzily 1 or 0, fp64
    // division by the root, narrowed per component).
    std::vector<float4> planes;
    planes.reserve((size_t)n_faces * 6);
    for (uint32_t f = 0; f < n_faces; ++f) {
        float nx = faces[f].normal[0] * -1.0f, ny = faces[f].normal[1] * -1.0f, nz = faces[f].normal[2] * -1.0f;
        double len = (double)nx * (double)nx + (double)ny * (double)ny + (double)nz * (double)nz;
        if (!(std::fabs(len - 1.0f) <= 0.000000000001 || std::fabs(len) <= 0.000000000001)) {
            len = std::sqrt(len);
            nx = (float)((double)nx / len);
            ny = (float)((double)ny / len);
            nz = (float)((double)nz / len);
        }
        const float *verts[3] = {faces[f].v0, faces[f].v1, faces[f].v2};
        for (const float *v : verts) {
            planes.push_back(make_float4(nx, ny, nz, 0.f));
            planes.push_back(make_float4(v[0], v[1], v[2], 0.f));
        }
    }
    CUDA_TRY(c, dalloc(&c->d_mesh_planes, planes.size()));
    CUDA_TRY(c, cudaMemcpy(c->d_mesh_planes, planes.data(), planes.size() * sizeof(float4), cudaMemcpyHostToDevice));
    c->P.mesh_planes = c->d_mesh_planes;
    c->P.mesh_plane_count = (int)(planes.size() / 2);
    return SPH_OK;
}

int sph_pin_host_buffer(sph_context *c, void *ptr, size_t bytes) {
    REQUIRE(c, c && ptr, SPH_ERR_ARGUMENT, "sph_pin_host_buffer: NULL argument");
    if (c->pinned_ptr) {
        cudaHostUnregister(c->pinned_ptr);
        c->pinned_ptr = nullptr;
    }
    CUDA_TRY(c, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    c->pinned_ptr = ptr;
    return SPH_OK;
}

// ---------------------------------------------------------------- phases
int sph_update_grid(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    CUDA_TRY(c, cudaSetDevice(c->device));
    PhaseTimer t(c, ms);
    enqueue_grid(c);
    c->s_valid = c->grid_valid = true;
    c->a_aligned = c->density_valid = c->forces_valid = false;
    int rc = check_launch(c, "update_grid");
    return rc ? rc : t.finish();
}

int sph_density_pressure(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, c->grid_valid && !c->a_aligned, SPH_ERR_STATE, "sph_density_pressure: call sph_update_grid first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    PhaseTimer t(c, ms);
    enqueue_density(c);
    c->density_valid = true;
    int rc = check_launch(c, "density_pressure");
    return rc ? rc : t.finish();
}

int sph_forces(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, c->grid_valid && c->density_valid && !c->a_aligned, SPH_ERR_STATE,
            "sph_forces: call sph_density_pressure first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    PhaseTimer t(c, ms);
    enqueue_forces(c);
    c->forces_valid = true;
    int rc = check_launch(c, "forces");
    return rc ? rc : t.finish();
}

int sph_collisions(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    if (ms) *ms = 0.0;  // fused into sph_integrate; the CPU path also reports 0 here
    return SPH_OK;
}

int sph_integrate(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, c->s_valid && c->forces_valid && !c->a_aligned, SPH_ERR_STATE, "sph_integrate: call sph_forces first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    PhaseTimer t(c, ms);
    enqueue_integrate(c);
    c->a_aligned = true;
    int rc = check_launch(c, "integrate");
    return rc ? rc : t.finish();
}

int sph_step(sph_context *c, int n_steps, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, n_steps >= 0, SPH_ERR_ARGUMENT, "sph_step: negative step count");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->slab) return slab_step(c, n_steps, ms);
    if (n_steps == 0 || (c->n == 0 && !c->emit_templates)) {
        if (ms) *ms = 0.0;
        return SPH_OK;
    }
    // Fountain still filling: every step starts with a device-side emission, the particle count changes, so the steps
    // are launched directly (fused kernels, no graph, no host transfer) until the emitter has run dry.
    if (c->emit_templates) {
        PhaseTimer t(c, ms);
        int k = 0;
        for (; k < n_steps; ++k) {
            if (enqueue_emit(c) == 0 && c->n > 0) break;  // cap reached: the rest runs on the stable-count path below
            if (c->n == 0) continue;
            enqueue_step(c);
            c->last_step_n = c->n;
        }
        c->steps += (uint64_t)k;
        if (k > 0 && c->n > 0) c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
        int rc = check_launch(c, "step (emitting)");
        if (rc) return rc;
        if (k == n_steps) return t.finish();
        double ms_head = 0.0, ms_tail = 0.0;
        if (ms) {
            rc = t.finish();
            if (rc) return rc;
            ms_head = *ms;
        }
        const uint32_t tpl = c->emit_templates;
        c->emit_templates = 0;  // dry: recurse once into the ordinary path for the remaining steps
        rc = sph_step(c, n_steps - k, ms ? &ms_tail : nullptr);
        c->emit_templates = tpl;
        if (ms) *ms = ms_head + ms_tail;
        return rc;
    }
    // Capture the step once the particle count is stable.
    if (c->opt_use_graph && !c->graph_exec && c->last_step_n == c->n) {
        cudaGraph_t graph = nullptr;
        const uint64_t launches_before = c->kernel_launches;
        CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_step(c);
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        c->kernel_launches = launches_before;
        if (e != cudaSuccess) return fail(c, SPH_ERR_CUDA, std::string("CUDA::") + cudaGetErrorName(e) + " in graph capture");
        e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(c, SPH_ERR_CUDA, std::string("CUDA::") + cudaGetErrorName(e) + " in graph instantiate");
        c->graph_n = c->n;
    }
    if (c->graph_exec && c->graph_n != c->n) drop_graph(c);
    c->last_step_n = c->n;

    if (c->opt_flush_l2) {
        // Every step is preceded by a write + read-back of a 256 MiB scratch buffer (> 126 MB L2; the read-back
        // leaves only clean scratch lines, so no write-back of scratch is charged to the step) and timed by its own
        // event pair; *ms is the sum of the per-step device times, the flushes are outside the timed spans.
        if (!c->d_flush) {
            c->flush_count = ((size_t)256 << 20) / sizeof(float4);
            CUDA_TRY(c, dalloc(&c->d_flush, c->flush_count));
        }
        while (c->step_events.size() < (size_t)(2 * n_steps)) {
            cudaEvent_t e;
            CUDA_TRY(c, cudaEventCreate(&e));
            c->step_events.push_back(e);
        }
        for (int k = 0; k < n_steps; ++k) {
            launch_flush_l2(c->d_flush, c->flush_count, c->stream);
            CUDA_TRY(c, cudaEventRecord(c->step_events[2 * k], c->stream));
            if (c->graph_exec) {
                CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->stream));
                c->graph_launches += 1;
                c->kernel_launches += kKernelsPerStep + ((use_mask_passes(c) && !c->opt_fuse_integrate) ? 1 : 0);
            } else {
                enqueue_step(c);
            }
            CUDA_TRY(c, cudaEventRecord(c->step_events[2 * k + 1], c->stream));
        }
        c->steps += (uint64_t)n_steps;
        c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
        int rc0 = check_launch(c, "step");
        if (rc0) return rc0;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        double total = 0.0;
        for (int k = 0; k < n_steps; ++k) {
            float t = 0.f;
            cudaEventElapsedTime(&t, c->step_events[2 * k], c->step_events[2 * k + 1]);
            total += (double)t;
        }
        if (ms) *ms = total;
        return SPH_OK;
    }

    PhaseTimer t(c, ms);
    for (int k = 0; k < n_steps; ++k) {
        if (c->graph_exec) {
            CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->stream));
            c->graph_launches += 1;
            c->kernel_launches += kKernelsPerStep + ((use_mask_passes(c) && !c->opt_fuse_integrate) ? 1 : 0);
        } else {
            enqueue_step(c);
        }
    }
    c->steps += (uint64_t)n_steps;
    c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
    int rc = check_launch(c, "step");
    return rc ? rc : t.finish();
}

// The same steps with CUDA events between the kernel groups (direct launches on the context's stream, L2 evicted
// before every step when the flush_l2 option is on): phase_ms[0..2] = grid build, density pass, forces + walls +
// integration summed over the steps, phase_ms[3] = sum of the whole-step spans.  For the bench's per-kernel split:
// the numbers come from the very steps that are being timed, not from a separate phase-by-phase run.
int sph_step_profiled(sph_context *c, int n_steps, double *phase_ms) {
    REQUIRE(c, c && phase_ms, SPH_ERR_ARGUMENT, "sph_step_profiled: NULL argument");
    REQUIRE(c, n_steps >= 0, SPH_ERR_ARGUMENT, "sph_step_profiled: negative step count");
    REQUIRE(c, !c->slab && !c->emit_templates, SPH_ERR_STATE, "sph_step_profiled: plain contexts only (no slab mode, no emitter)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (int k = 0; k < 4; ++k) phase_ms[k] = 0.0;
    if (n_steps == 0 || c->n == 0) return SPH_OK;
    if (c->opt_flush_l2 && !c->d_flush) {
        c->flush_count = ((size_t)256 << 20) / sizeof(float4);
        CUDA_TRY(c, dalloc(&c->d_flush, c->flush_count));
    }
    while (c->step_events.size() < (size_t)(4 * n_steps)) {
        cudaEvent_t e;
        CUDA_TRY(c, cudaEventCreate(&e));
        c->step_events.push_back(e);
    }
    const bool fused = use_mask_passes(c) && c->opt_fuse_integrate;
    for (int k = 0; k < n_steps; ++k) {
        if (c->opt_flush_l2) launch_flush_l2(c->d_flush, c->flush_count, c->stream);
        CUDA_TRY(c, cudaEventRecord(c->step_events[4 * k + 0], c->stream));
        enqueue_grid(c);
        CUDA_TRY(c, cudaEventRecord(c->step_events[4 * k + 1], c->stream));
        enqueue_density(c);
        CUDA_TRY(c, cudaEventRecord(c->step_events[4 * k + 2], c->stream));
        if (fused) {
            launch_forces_mask(c->nb, c->dp, c->nb_count, c->g.key_s, c->g.cell_start, c->acc, 0, (int)c->n, c->P, c->stream,
                               c->pos_s, c->pos_a, c->vel_a);
            c->kernel_launches += 2;
            c->in_off = 0;
        } else {
            enqueue_forces(c);
            enqueue_integrate(c);
        }
        CUDA_TRY(c, cudaEventRecord(c->step_events[4 * k + 3], c->stream));
    }
    c->steps += (uint64_t)n_steps;
    c->last_step_n = c->n;
    c->s_valid = c->grid_valid = c->density_valid = c->forces_valid = c->a_aligned = true;
    if (int rc = check_launch(c, "step (profiled)")) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n_steps; ++k) {
        float t = 0.f;
        for (int ph = 0; ph < 3; ++ph) {
            cudaEventElapsedTime(&t, c->step_events[4 * k + ph], c->step_events[4 * k + ph + 1]);
            phase_ms[ph] += (double)t;
        }
        cudaEventElapsedTime(&t, c->step_events[4 * k], c->step_events[4 * k + 3]);
        phase_ms[3] += (double)t;
    }
    return SPH_OK;
}

int sph_synchronize(sph_context *c) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPH_OK;
}

// ---------------------------------------------------------------- taps
int sph_download_keys(sph_context *c, int32_t *keys) {
    REQUIRE(c, c && keys, SPH_ERR_ARGUMENT, "sph_download_keys: NULL argument");
    REQUIRE(c, c->grid_valid, SPH_ERR_STATE, "sph_download_keys: grid not built");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_keys: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    launch_scatter_cell_ids(c->pos_s, c->g.key_s, c->d_tmp_i32, (int)c->n, c->P, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(keys, c->d_tmp_i32, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (int rc = id_precondition(c, "sph_download_keys")) return rc;
    return check_launch(c, "download_keys");
}

int sph_download_permutation(sph_context *c, uint32_t *sorted_ids) {
    REQUIRE(c, c && sorted_ids, SPH_ERR_ARGUMENT, "sph_download_permutation: NULL argument");
    REQUIRE(c, c->grid_valid, SPH_ERR_STATE, "sph_download_permutation: grid not built");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_permutation: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    // the device order is (cell, x bin, id); the tap reports the order stable by (cell id, particle id)
    launch_cell_id_permutation(c->pos_s, c->g.key_s, c->g.cell_start, reinterpret_cast<unsigned *>(c->d_tmp_i32), (int)c->n,
                               c->P, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(sorted_ids, c->d_tmp_i32, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return check_launch(c, "download_permutation");
}

int sph_download_cell_start(sph_context *c, int32_t *cell_start) {
    REQUIRE(c, c && cell_start, SPH_ERR_ARGUMENT, "sph_download_cell_start: NULL argument");
    REQUIRE(c, c->grid_valid, SPH_ERR_STATE, "sph_download_cell_start: grid not built");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int *d_coarse = nullptr;  // one entry per reference cell out of the finer sort-key scan
    CUDA_TRY(c, dalloc(&d_coarse, (size_t)c->P.n_cells + 1));
    launch_coarse_cell_start(c->g.cell_start, d_coarse, c->P.n_cells, c->P.xb, c->stream);
    c->kernel_launches += 1;
    cudaError_t e = cudaMemcpyAsync(cell_start, d_coarse, ((size_t)c->P.n_cells + 1) * sizeof(int), cudaMemcpyDeviceToHost,
                                    c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_coarse);
    CUDA_TRY(c, e);
    return check_launch(c, "download_cell_start");
}

int sph_download_density_pressure_accel(sph_context *c, float *density, float *pressure, float *accel3) {
    REQUIRE(c, c && density && pressure && accel3, SPH_ERR_ARGUMENT, "sph_download_density_pressure_accel: NULL argument");
    REQUIRE(c, c->s_valid && c->density_valid, SPH_ERR_STATE, "sph_download_density_pressure_accel: nothing computed yet");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_density_pressure_accel: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    float *rho = c->d_tmp_f32, *prs = rho + n, *a3 = prs + n;
    launch_scatter_dpa_by_id(c->pos_s, c->dp, c->acc, rho, prs, a3, (int)n, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(density, rho, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(pressure, prs, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(accel3, a3, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (int rc = id_precondition(c, "sph_download_density_pressure_accel")) return rc;
    return check_launch(c, "download_density_pressure_accel");
}

int sph_download_neighbours(sph_context *c, int32_t *counts, int32_t *lists, uint64_t lists_capacity, uint64_t *total) {
    REQUIRE(c, c && counts, SPH_ERR_ARGUMENT, "sph_download_neighbours: NULL argument");
    REQUIRE(c, c->grid_valid && c->density_valid, SPH_ERR_STATE, "sph_download_neighbours: run sph_density_pressure first");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_neighbours: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    launch_scatter_by_id_i32(c->pos_s, c->nb_count, c->d_tmp_i32, (int)n, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(counts, c->d_tmp_i32, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (int rc = id_precondition(c, "sph_download_neighbours")) return rc;
    std::vector<long long> offsets(n + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + counts[i];
    if (total) *total = (uint64_t)offsets[n];
    if (!lists) return check_launch(c, "download_neighbours");
    REQUIRE(c, lists_capacity >= (uint64_t)offsets[n], SPH_ERR_ARGUMENT, "sph_download_neighbours: lists buffer too small");
    long long *d_off = nullptr;
    int *d_lists = nullptr;
    CUDA_TRY(c, dalloc(&d_off, n + 1));
    cudaError_t e = dalloc(&d_lists, (size_t)offsets[n]);
    if (e != cudaSuccess) {
        cudaFree(d_off);
        CUDA_TRY(c, e);
    }
    cudaMemcpyAsync(d_off, offsets.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    launch_neighbour_lists(c->pos_s, c->g.key_s, c->g.cell_start, d_off, d_lists, (int)n, c->P, c->stream);
    c->kernel_launches += 1;
    cudaMemcpyAsync(lists, d_lists, (size_t)offsets[n] * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    cudaFree(d_off);
    cudaFree(d_lists);
    CUDA_TRY(c, e);
    for (size_t i = 0; i < n; ++i) std::sort(lists + offsets[i], lists + offsets[i + 1]);
    return check_launch(c, "download_neighbours");
}

int sph_download_mask_neighbours(sph_context *c, int32_t *counts, int32_t *lists, uint64_t lists_capacity, uint64_t *total) {
    REQUIRE(c, c && counts, SPH_ERR_ARGUMENT, "sph_download_mask_neighbours: NULL argument");
    REQUIRE(c, c->grid_valid && c->density_valid, SPH_ERR_STATE, "sph_download_mask_neighbours: run sph_density_pressure first");
    REQUIRE(c, use_mask_passes(c), SPH_ERR_STATE, "sph_download_mask_neighbours: the bitmask passes are not in use (variant 0 or a grid narrower than 4 cells)");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_download_mask_neighbours: indexed by global id - not available in slab mode");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    launch_mask_lists(c->nb, c->pos_s, c->g.key_s, c->g.cell_start, nullptr, nullptr, c->d_tmp_i32, (int)n, c->P, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(counts, c->d_tmp_i32, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (int rc = id_precondition(c, "sph_download_mask_neighbours")) return rc;
    std::vector<long long> offsets(n + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + counts[i];
    if (total) *total = (uint64_t)offsets[n];
    if (!lists) return check_launch(c, "download_mask_neighbours");
    REQUIRE(c, lists_capacity >= (uint64_t)offsets[n], SPH_ERR_ARGUMENT, "sph_download_mask_neighbours: lists buffer too small");
    long long *d_off = nullptr;
    int *d_lists = nullptr;
    CUDA_TRY(c, dalloc(&d_off, n + 1));
    cudaError_t e = dalloc(&d_lists, (size_t)offsets[n]);
    if (e != cudaSuccess) {
        cudaFree(d_off);
        CUDA_TRY(c, e);
    }
    cudaMemcpyAsync(d_off, offsets.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    launch_mask_lists(c->nb, c->pos_s, c->g.key_s, c->g.cell_start, d_off, d_lists, nullptr, (int)n, c->P, c->stream);
    c->kernel_launches += 1;
    cudaMemcpyAsync(lists, d_lists, (size_t)offsets[n] * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    cudaFree(d_off);
    cudaFree(d_lists);
    CUDA_TRY(c, e);
    for (size_t i = 0; i < n; ++i) std::sort(lists + offsets[i], lists + offsets[i + 1]);
    return check_launch(c, "download_mask_neighbours");
}

// ---------------------------------------------------------------- all-pairs variant
static int brute_snapshot(sph_context *c) {
    // S := A (no sort): the aux arrays are then aligned with the input order
    if (!c->s_valid || c->a_aligned) {
        CUDA_TRY(c, cudaMemcpyAsync(c->pos_s, c->pos_a, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->vel_s, c->vel_a, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        c->s_valid = true;
        c->grid_valid = false;
        c->a_aligned = false;
    }
    return SPH_OK;
}

int sph_brute_density_pressure(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = brute_snapshot(c);
    if (rc) return rc;
    PhaseTimer t(c, ms);
    launch_brute_density(c->pos_s, c->dp, c->nb_count, (int)c->n, c->P, c->stream);
    c->kernel_launches += 1;
    c->density_valid = true;
    c->forces_valid = false;
    rc = check_launch(c, "brute_density_pressure");
    return rc ? rc : t.finish();
}

int sph_brute_forces(sph_context *c, double *ms) {
    REQUIRE(c, c, SPH_ERR_ARGUMENT, "NULL context");
    REQUIRE(c, c->s_valid && c->density_valid && !c->a_aligned, SPH_ERR_STATE, "sph_brute_forces: density first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    PhaseTimer t(c, ms);
    launch_brute_forces(c->pos_s, c->vel_s, c->dp, c->acc, (int)c->n, c->P, c->stream);
    c->kernel_launches += 1;
    c->forces_valid = true;
    int rc = check_launch(c, "brute_forces");
    return rc ? rc : t.finish();
}

int sph_brute_neighbour_counts(sph_context *c, int32_t *counts) {
    REQUIRE(c, c && counts, SPH_ERR_ARGUMENT, "sph_brute_neighbour_counts: NULL argument");
    REQUIRE(c, c->s_valid && c->density_valid, SPH_ERR_STATE, "sph_brute_neighbour_counts: density first");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_brute_neighbour_counts: indexed by global id - not available in slab mode (use sph_download_owned)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    launch_scatter_by_id_i32(c->pos_s, c->nb_count, c->d_tmp_i32, (int)c->n, c->stream);
    c->kernel_launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(counts, c->d_tmp_i32, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (int rc = id_precondition(c, "sph_brute_neighbour_counts")) return rc;
    return check_launch(c, "brute_neighbour_counts");
}

// ---------------------------------------------------------------- statistics
int sph_stats(sph_context *c, double *out6) {
    REQUIRE(c, c && out6, SPH_ERR_ARGUMENT, "sph_stats: NULL argument");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t so = c->slab ? c->in_off : 0;
    const uint32_t sn = c->slab ? c->slab->n_own : c->n;
    launch_stats(view_pos(c) + so, view_vel(c) + so, (int)sn, c->d_stats, c->stream);
    c->kernel_launches += 1;
    double h[8];
    CUDA_TRY(c, cudaMemcpyAsync(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const double n = (double)sn;
    unsigned bits;
    std::memcpy(&bits, &h[5], sizeof(bits));
    bits = (bits & 0x80000000u) ? (bits & 0x7fffffffu) : ~bits;
    float ymax;
    std::memcpy(&ymax, &bits, sizeof(ymax));
    if (c->slab) {
        // partial sums of this rank: the caller adds them over ranks (and takes the max of out6[4])
        out6[0] = 0.5 * (double)c->P.mass * h[0];
        out6[1] = h[1]; out6[2] = h[2]; out6[3] = h[3];
        out6[4] = n > 0 ? (double)ymax + (double)c->cfg.box[1] / 2.0 : 0;
        out6[5] = h[4];
        return check_launch(c, "stats");
    }
    out6[0] = 0.5 * (double)c->P.mass * h[0];
    out6[1] = n > 0 ? h[1] / n : 0;
    out6[2] = n > 0 ? h[2] / n : 0;
    out6[3] = n > 0 ? h[3] / n : 0;
    out6[4] = n > 0 ? (double)ymax + (double)c->cfg.box[1] / 2.0 : 0;
    out6[5] = n > 0 ? h[4] / n : 0;
    return check_launch(c, "stats");
}

// Order statistic of the fill height y + b/2 by a min/max pass and two histogram passes: 4096 bins over [ymin, ymax],
// then 4096 bins inside the bin that holds rank k = floor(q (n - 1)) (the definition oracle_stats uses, SURVEY.md
// §8c).  The answer is the centre of a sub-bin of width (ymax - ymin) / 4096^2 — below fp32 resolution of y.
int sph_fill_height_percentile(sph_context *c, double q, double *height) {
    REQUIRE(c, c && height, SPH_ERR_ARGUMENT, "sph_fill_height_percentile: NULL argument");
    REQUIRE(c, q >= 0.0 && q <= 1.0, SPH_ERR_ARGUMENT, "sph_fill_height_percentile: q outside [0, 1]");
    REQUIRE(c, !c->slab, SPH_ERR_STATE, "sph_fill_height_percentile: a global order statistic - not available in slab mode");
    CUDA_TRY(c, cudaSetDevice(c->device));
    *height = 0.0;
    if (c->n == 0) return SPH_OK;
    // range of the finite y values (particles may sit outside the box: the walls are penalty forces)
    launch_minmax_y(view_pos(c), (int)c->n, c->d_hist, c->stream);
    c->kernel_launches += 1;
    unsigned mm[2];
    CUDA_TRY(c, cudaMemcpyAsync(mm, c->d_hist, sizeof(mm), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    REQUIRE(c, mm[0] <= mm[1], SPH_ERR_STATE, "sph_fill_height_percentile: no particle has a finite y coordinate");
    auto from_ordered = [](unsigned u) {
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        float f;
        std::memcpy(&f, &u, sizeof(f));
        return (double)f;
    };
    const double ymin = from_ordered(mm[0]), ymax = from_ordered(mm[1]);
    const uint64_t k = (uint64_t)(q * (double)(c->n - 1));
    std::vector<unsigned> h(kHistBins);
    double lo = ymin, width = std::max((ymax - ymin) * (1.0 + 1e-9), 1e-30) / kHistBins;
    uint64_t below = 0;  // particles in bins before the selected one
    for (int pass = 0; pass < 2; ++pass) {
        launch_hist_y(view_pos(c), (int)c->n, lo, 1.0 / width, pass == 0 ? 1 : 0, c->d_hist, c->stream);
        c->kernel_launches += 1;
        CUDA_TRY(c, cudaMemcpyAsync(h.data(), c->d_hist, kHistBins * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        // pass 0 puts non-finite values into bin 0 (clamped); pass 1 only counts particles of the selected bin
        uint64_t run = pass == 0 ? 0 : below;
        int b = 0;
        for (; b < kHistBins - 1; ++b) {
            if (run + h[b] > k) break;
            run += h[b];
        }
        below = run;
        lo += width * b;
        if (pass == 0) width /= kHistBins;
    }
    *height = lo + 0.5 * width + (double)c->cfg.box[1] / 2.0;
    return check_launch(c, "fill_height_percentile");
}

// ---------------------------------------------------------------- options / counters
int sph_set_option(sph_context *c, const char *name, int value) {
    REQUIRE(c, c && name, SPH_ERR_ARGUMENT, "sph_set_option: NULL argument");
    const std::string k(name);
    if (k == "neighbour_variant") c->opt_neighbour_variant = value;
    else if (k == "use_graph") c->opt_use_graph = value;
    else if (k == "fuse_integrate") c->opt_fuse_integrate = value;
    else if (k == "count_neighbours") c->opt_count_neighbours = value;
    else if (k == "flush_l2") { c->opt_flush_l2 = value; return SPH_OK; }
    else if (k == "tuning") c->P.tuning = value;
    else if (k == "slab_overlap") {
        REQUIRE(c, c->slab, SPH_ERR_STATE, "sph_set_option: slab_overlap needs a slab context");
        c->slab->opt_overlap = value;
        return SPH_OK;
    }
    else if (k == "slab_rebalance") {  // 0 static faces, 1 faces follow the load (default), 2 deterministic test pattern
        REQUIRE(c, c->slab, SPH_ERR_STATE, "sph_set_option: slab_rebalance needs a slab context");
        c->slab->opt_rebalance = value;
        return SPH_OK;
    }
    else return fail(c, SPH_ERR_ARGUMENT, "sph_set_option: unknown option " + k);
    drop_graph(c);
    return SPH_OK;
}

int sph_get_counter(const sph_context *c, const char *name, uint64_t *value) {
    if (!c || !name || !value) return SPH_ERR_ARGUMENT;
    const std::string k(name);
    if (k == "kernel_launches") *value = c->kernel_launches;
    else if (k == "graph_launches") *value = c->graph_launches;
    else if (k == "steps") *value = c->steps;
    else if (k == "slab_face_moves") *value = c->slab ? c->slab->face_moves : 0;
    else if (k == "slab_load_us") *value = c->slab ? (uint64_t)c->slab->load_us : 0;
    else if (k == "neighbour_pairs") {  // sum of the per-particle neighbour counts of the last density pass (self included)
        if (cudaSetDevice(c->device) != cudaSuccess) return SPH_ERR_CUDA;
        sph_context *m = const_cast<sph_context *>(c);
        launch_sum_i32(c->nb_count, (int)c->n, reinterpret_cast<unsigned long long *>(c->d_stats), c->stream);
        m->kernel_launches += 1;
        unsigned long long v = 0;
        if (cudaMemcpyAsync(&v, c->d_stats, sizeof(v), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess)
            return SPH_ERR_CUDA;
        *value = (uint64_t)v;
    }
    else if (k == "overflow_particles" || k == "slab_far_movers") {
        // overflow_particles: particles of the last density pass whose hit words did not fit;
        // slab_far_movers: particles the boundary-only exchange would have missed (must stay 0).
        // Device counters: read on the context's device, ordered behind the work already queued on its streams.
        int v = 0;
        const int *src = k == "overflow_particles" ? c->nb.ovf : (c->slab ? c->slab->d_counters + 4 : nullptr);
        if (src) {
            if (cudaSetDevice(c->device) != cudaSuccess) return SPH_ERR_CUDA;
            if (c->slab && c->slab->comm_stream && cudaStreamSynchronize(c->slab->comm_stream) != cudaSuccess) return SPH_ERR_CUDA;
            if (cudaMemcpyAsync(&v, src, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess)
                return SPH_ERR_CUDA;
        }
        *value = (uint64_t)v;
    }
    else return SPH_ERR_ARGUMENT;
    return SPH_OK;
}

int sph_comm_unique_id(uint8_t out[128]) {
    REQUIRE(nullptr, out, SPH_ERR_ARGUMENT, "sph_comm_unique_id: NULL buffer");
    std::string why;
    NcclApi *api = nccl_api(&why);
    if (!api) return fail(nullptr, SPH_ERR_COMM, why);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(nullptr, api, api->GetUniqueId(&id));
    std::memcpy(out, &id, 128);
    return SPH_OK;
}

int sph_comm_local_id(int32_t world, uint8_t out[128]) {
    REQUIRE(nullptr, out && world >= 1, SPH_ERR_ARGUMENT, "sph_comm_local_id: bad argument");
    LocalComm *lc = new LocalComm();
    lc->world = world;
    lc->up.reset(new LocalEdge[world]);
    lc->down.reset(new LocalEdge[world]);
    std::memset(out, 0, 128);
    std::memcpy(out, kLocalMagic, sizeof(kLocalMagic));
    std::memcpy(out + sizeof(kLocalMagic), &lc, sizeof(lc));
    return SPH_OK;  // the mailbox is freed with the last slab context created from this id
}

int sph_slab_plan(int32_t rz, int32_t world, int32_t rank, int32_t *z0, int32_t *z1) {
    REQUIRE(nullptr, z0 && z1 && world >= 1 && rank >= 0 && rank < world, SPH_ERR_ARGUMENT, "sph_slab_plan: bad argument");
    REQUIRE(nullptr, rz / world >= 4, SPH_ERR_ARGUMENT, "sph_slab_plan: every slab needs at least 4 z-layers");
    int a, b;
    slab_plan(rz, world, rank, &a, &b);
    *z0 = a;
    *z1 = b;
    return SPH_OK;
}

int sph_slab_face_shift(const int32_t lower4[4], const int32_t upper4[4], int32_t shift, int32_t shift_max, int32_t mode,
                        uint64_t exchange, int32_t face) {
    if (!lower4 || !upper4) return 0;
    const int lo[4] = {lower4[0], lower4[1], lower4[2], lower4[3]}, hi[4] = {upper4[0], upper4[1], upper4[2], upper4[3]};
    return face_shift(lo, hi, shift, shift_max, mode, exchange, face);
}

int sph_slab_create(const sph_config *cfg, sph_context **out) {
    REQUIRE(nullptr, cfg && out, SPH_ERR_ARGUMENT, "sph_slab_create: NULL argument");
    *out = nullptr;
    REQUIRE(nullptr, cfg->world >= 1 && cfg->rank >= 0 && cfg->rank < cfg->world, SPH_ERR_ARGUMENT, "sph_slab_create: bad rank/world");
    REQUIRE(nullptr, cfg->grid_res[2] / cfg->world >= 4, SPH_ERR_ARGUMENT, "sph_slab_create: every slab needs at least 4 z-layers");
    int z0, z1;
    slab_plan(cfg->grid_res[2], cfg->world, cfg->rank, &z0, &z1);
    // the faces may move by up to shift_max layers from this plan (load balance, see Slab): the local grid gets that many
    // spare layers on either side, so that z_base stays put
    const int shift_max = cfg->world > 1 ? std::min(std::max((z1 - z0) / 5, 3), 128) : 0;
    const int z_base = std::max(z0 - 2 - shift_max, 0), z_top = std::min(z1 + 2 + shift_max, cfg->grid_res[2]);
    sph_config local = *cfg;
    local.grid_res[2] = z_top - z_base;  // owned layers + two ghost layers per interior face + room for the faces to move
    local.world = 1;
    int rc = sph_create(&local, out);
    if (rc) return rc;
    sph_context *c = *out;
    c->cfg = *cfg;
    c->P.rz_global = cfg->grid_res[2];
    c->P.z_base = z_base;
    c->P.own_z0 = c->P.next_z0 = z0;
    c->P.own_z1 = c->P.next_z1 = z1;
    c->opt_use_graph = 0;  // the particle count changes every step
    Slab *s = new Slab();
    c->slab = s;
    s->rank = cfg->rank;
    s->world = cfg->world;
    s->z0 = s->z0_init = z0;
    s->z1 = s->z1_init = z1;
    s->shift_max = shift_max;
    s->z_base = z_base;
    s->rz_local = z_top - z_base;
    slab_set_layer_windows(c);
    // face buffers: up to 4 layers (2 ghost + 2 of slack for migrants) at 48 particles per cell
    const size_t face = (size_t)cfg->grid_res[0] * cfg->grid_res[1] * 4 * 48;
    s->cap_face = (int)std::min<size_t>(face, c->cap);
#define SLAB_TRY(call)                                                             \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            std::string m = std::string("CUDA::") + cudaGetErrorName(e_) + " | " #call; \
            sph_destroy(c);                                                        \
            *out = nullptr;                                                        \
            return fail(nullptr, SPH_ERR_CUDA, m);                                 \
        }                                                                          \
    } while (0)
    SLAB_TRY(dalloc(&s->down_pos, (size_t)s->cap_face));
    SLAB_TRY(dalloc(&s->down_vel, (size_t)s->cap_face));
    SLAB_TRY(dalloc(&s->up_pos, (size_t)s->cap_face));
    SLAB_TRY(dalloc(&s->up_vel, (size_t)s->cap_face));
    SLAB_TRY(dalloc(&s->d_counters, (size_t)8));  // [0..3] exchange counts, [4] particles that moved > 2 layers
    SLAB_TRY(cudaMemset(s->d_counters, 0, 8 * sizeof(int)));
    SLAB_TRY(cudaMallocHost(reinterpret_cast<void **>(&s->h_pinned), 32 * sizeof(int)));
    SLAB_TRY(dalloc(&s->d_aux, (size_t)12));
    SLAB_TRY(cudaMemset(s->d_aux, 0, 12 * sizeof(int)));
    for (auto &pair : s->ev_busy)
        for (cudaEvent_t &e : pair) SLAB_TRY(cudaEventCreate(&e));
    // highest priority: the boundary kernels, the pack and NCCL must get SM slots ahead of the interior force kernel
    // that is already filling the device on the compute stream
    int prio_least = 0, prio_greatest = 0;
    SLAB_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    SLAB_TRY(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, prio_greatest));
    SLAB_TRY(cudaEventCreateWithFlags(&s->ev_ranges, cudaEventDisableTiming));
    SLAB_TRY(cudaEventCreateWithFlags(&s->ev_boundary, cudaEventDisableTiming));
    SLAB_TRY(cudaEventCreateWithFlags(&s->ev_counts, cudaEventDisableTiming));
    SLAB_TRY(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
#undef SLAB_TRY
    if (cfg->world > 1 && std::memcmp(cfg->nccl_id, kLocalMagic, sizeof(kLocalMagic)) == 0) {
        // loop-back transport: the id carries the address of the mailbox shared by the ranks of this process
        LocalComm *lc = nullptr;
        std::memcpy(&lc, cfg->nccl_id + sizeof(kLocalMagic), sizeof(lc));
        if (!lc || lc->world != cfg->world) {
            sph_destroy(c);
            *out = nullptr;
            return fail(nullptr, SPH_ERR_ARGUMENT, "sph_slab_create: loop-back id does not match world");
        }
        s->local = lc;
        lc->refs.fetch_add(1);
        // this rank creates the events of the edges it SENDS on (they are recorded on its streams)
        for (LocalEdge *e : {s->rank + 1 < s->world ? &lc->up[s->rank] : nullptr, s->rank > 0 ? &lc->down[s->rank] : nullptr})
            if (e) {
                std::lock_guard<std::mutex> g(e->m);
                if (!e->packed) cudaEventCreateWithFlags(&e->packed, cudaEventDisableTiming);
                if (!e->copied) cudaEventCreateWithFlags(&e->copied, cudaEventDisableTiming);
            }
    } else if (cfg->world > 1) {
        std::string why;
        NcclApi *api = nccl_api(&why);
        if (!api) {
            sph_destroy(c);
            *out = nullptr;
            return fail(nullptr, SPH_ERR_COMM, why);
        }
        ncclUniqueId id;
        std::memcpy(&id, cfg->nccl_id, 128);
        ncclResult_t r = api->CommInitRank(&s->comm, cfg->world, id, cfg->rank);
        if (r != ncclSuccess) {
            std::string m = std::string("NCCL::") + api->GetErrorString(r) + " | ncclCommInitRank";
            sph_destroy(c);
            *out = nullptr;
            return fail(nullptr, SPH_ERR_COMM, m);
        }
    }
    return SPH_OK;
}

int sph_slab_info(const sph_context *c, int32_t out8[8]) {
    if (!c || !out8 || !c->slab) return SPH_ERR_ARGUMENT;
    const Slab &s = *c->slab;
    out8[0] = s.rank; out8[1] = s.world; out8[2] = s.z0; out8[3] = s.z1; out8[4] = s.z_base; out8[5] = s.rz_local;
    out8[6] = (int32_t)s.n_own; out8[7] = (int32_t)(s.exchanges ? s.sent_particles / s.exchanges : 0);
    return SPH_OK;
}

int sph_download_owned(sph_context *c, sph_particle *aos, uint32_t capacity, uint32_t *n_out) {
    REQUIRE(c, c && aos, SPH_ERR_ARGUMENT, "sph_download_owned: NULL argument");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const uint32_t n = c->slab ? c->slab->n_own : c->n, off = c->slab ? c->in_off : 0;
    REQUIRE(c, capacity >= n, SPH_ERR_ARGUMENT, "sph_download_owned: buffer too small");
    const bool aux = aux_aligned(c) && c->a_aligned;  // aux arrays share A's order after a full step
    for (uint32_t base = 0; base < n;) {
        const uint32_t chunk = (uint32_t)std::min<size_t>(n - base, c->stage_cap);
        const size_t o = (size_t)off + base;
        launch_soa_to_aos((c->a_aligned || !c->s_valid ? c->pos_a : c->pos_s) + o, (c->a_aligned || !c->s_valid ? c->vel_a : c->vel_s) + o,
                          (aux && c->forces_valid) ? c->acc + o : nullptr, (aux && c->density_valid) ? c->dp + o : nullptr,
                          aux ? c->g.key_s + o : nullptr, c->d_stage,
                          -1, (int)chunk, (int)chunk, c->P, c->stream);
        c->kernel_launches += 1;
        CUDA_TRY(c, cudaMemcpyAsync(aos + base, c->d_stage, (size_t)chunk * sizeof(sph_particle), cudaMemcpyDeviceToHost,
                                    c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        base += chunk;
    }
    if (n_out) *n_out = n;
    return check_launch(c, "download_owned");
}

}  // extern "C"
