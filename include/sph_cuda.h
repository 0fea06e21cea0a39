/*
 * sph_cuda.h — C ABI of the B200-native SPH time step (libsph_cuda.so).
 *
 * This is the drop-in boundary.  It replaces the reference's device layer L1
 * (CLWrapper + CLPlatforms) and device code L0 (the .cl files in resources/kernels) for ONE path:
 * CBaseParticleSimulator::step() = updateGrid -> updateDensityPressure -> updateForces ->
 * updateCollisions -> integrate (src/CBaseParticleSimulator.cpp:116-144).  The only caller is the
 * C++ simulator CCUDAParticleSimulator (gmu-water-simulation_b200/core), which sits behind
 * CBaseParticleSimulator exactly like CCPUParticleSimulator / CGPUParticleSimulator do.
 *
 * Conventions: plain C types only; every function returns an int status (SPH_OK == 0);
 * sph_last_error() gives the message (≙ CLWrapper::getErrorMessage/checkError,
 * include/CLWrapper.h:35-36).  Thread-compatible, not thread-safe (the reference is strictly
 * single-threaded, src/CBaseParticleSimulator.cpp:35,76-81).  No OpenCL, no CPU fallback: when no
 * CUDA device is usable every entry point fails with SPH_ERR_CUDA.
 *
 * Numerics follow the reference's CPU path (src/CCPUParticleSimulator.cpp): cell keys use fp64 math
 * on fp32 positions, the neighbour predicate is the un-contracted fp32 (dx*dx+dy*dy)+dz*dz <= h*h,
 * so keys, the canonical (cell,id) permutation and neighbour sets are bit-exact; density, pressure
 * and acceleration are fp32 (rel 1e-5 vs. the mixed fp32/fp64 CPU path).
 *
 * All file:line citations are relative to the reference repository root.
 */
#ifndef SPH_CUDA_H
#define SPH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_ABI_VERSION 1

enum sph_status {
    SPH_OK = 0,
    SPH_ERR_CUDA = 1,      /* a CUDA runtime call failed (≙ CLException, include/CLWrapper.h:18-22) */
    SPH_ERR_ARGUMENT = 2,  /* bad argument / capacity exceeded */
    SPH_ERR_STATE = 3,     /* call order violated (e.g. forces before the grid was built) */
    SPH_ERR_COMM = 4       /* NCCL failure in slab mode */
};

/* CParticle::Physics / ParticleCL, byte for byte: 80 B, 16-aligned
 * (include/CParticle.h:19-43, resources/kernels/sph_common.cl:29-39).  float3/int3 are 16 B. */
typedef struct sph_particle {
    float position[4];
    float velocity[4];
    float acceleration[4];
    int32_t grid_position[4];
    float density;
    float pressure;
    uint32_t id;
    uint32_t cell_id;
} sph_particle;

/* sWall / WallCL: 32 B (include/CCollisionGeometry.h:23-30, resources/kernels/sph_common.cl:23-27) */
typedef struct sph_wall {
    float normal[4];
    float position[4];
} sph_wall;

/* Everything the reference passes to its kernels as arguments or __constant data
 * (CBaseParticleSimulator members include/CBaseParticleSimulator.h:80-90; constants
 * include/CParticle.h:80-84, include/CCollisionGeometry.h:20-21). */
typedef struct sph_config {
    float box[3];            /* m_boxSize; the reference only builds cubes, a non-cubic box is the tank extension */
    int32_t grid_res[3];     /* ceil(box/h), src/CBaseParticleSimulator.cpp:27-31 */
    float h;                 /* CParticle::h */
    float dt;                /* 0.01f, src/CBaseParticleSimulator.cpp:7 */
    float mass;              /* CParticle::mass */
    float viscosity;         /* CParticle::viscosity */
    float gas_stiffness;     /* CParticle::gas_stiffness */
    float rest_density;      /* CParticle::rest_density */
    float gravity[3];        /* (0, GRAVITY_ACCELERATION, 0), include/CBaseParticleSimulator.h:19 */
    double wall_k;           /* WALL_K 10000.0 — the reference's macros are double literals and enter the */
    double wall_damping;     /* WALL_DAMPING (-0.9)   arithmetic as doubles (src/CCollisionGeometry.cpp:124-128) */
    double wall_skin;        /* 0.01 "particle radius", src/CCollisionGeometry.cpp:124 */
    int32_t wall_count;      /* 6 */
    sph_wall walls[6];       /* left,bottom,back,right,top,front; include/CCollisionGeometry.h:79-120 */
    uint32_t max_particles;  /* m_maxParticlesCount: device capacity */
    int32_t device;          /* CUDA device ordinal (≙ the cl::Device ctor argument) */
    /* slab decomposition along z (multi-GPU extension; world == 1 means single device) */
    int32_t rank;
    int32_t world;
    uint8_t nccl_id[128];    /* ncclUniqueId from sph_comm_unique_id(), identical on all ranks */
} sph_config;

typedef struct sph_context sph_context;

/* ---- enumeration (≙ CLPlatforms::getAllPlatforms/getDevices/getDeviceInfo, include/CLPlatforms.h:10-15) ---- */
int sph_abi_version(void);
int sph_device_count(int *count);
int sph_device_name(int device, char *buf, size_t len);

/* ---- lifetime (≙ CLWrapper ctor + createBuffer, include/CLWrapper.h:53,64) ---- */
/* Fill cfg with the reference's constants, ceil(box/h) grid and the six +-box/2 walls. */
int sph_config_init(sph_config *cfg, float box_x, float box_y, float box_z, uint32_t max_particles);
int sph_create(const sph_config *cfg, sph_context **out);
int sph_destroy(sph_context *ctx);
const char *sph_last_error(const sph_context *ctx); /* ctx may be NULL: error of the last failed create/enumeration */

/* ---- state (≙ enqueueWrite/enqueueRead, include/CLWrapper.h:66-67) ---- */
/* Particle ids (sph_particle.id) must be unique and < max_particles; the read-backs "indexed by particle id" below
 * write host slot [id], so to fill every slot the ids must be a permutation of 0..n-1 — what setupScene /
 * generateParticles produce (src/CBaseParticleSimulator.cpp:69).  An id >= max_particles is detected on the device at
 * upload; the next by-id read-back then fails with SPH_ERR_ARGUMENT instead of writing out of bounds, and ids in
 * [n, max_particles) are skipped.  (Slab contexts carry global ids and only offer sph_download_owned.) */
/* Replace the device state with n particles from the 80-byte AoS host mirror (m_clParticles).  From a page-locked
 * buffer (sph_pin_host_buffer) the copy is asynchronous on the context's stream: leave the records alone until
 * the next synchronising call (a timed phase, sph_download_particles, sph_synchronize). */
int sph_upload_particles(sph_context *ctx, const sph_particle *aos, uint32_t n);
/* Fountain emission (src/CBaseParticleSimulator.cpp:187-210): append n_new particles. */
int sph_append_particles(sph_context *ctx, const sph_particle *aos, uint32_t n_new);
/* Device-side snapshot of the state (positions, velocities, their array order) in slot 0 or 1, and its exact
 * restoration: a replay then runs the same steps with the same memory layout (used by the bench to time the
 * per-kernel split and a fixed window on identical steps).  Call between steps. */
int sph_state_save(sph_context *ctx, int slot);
int sph_state_restore(sph_context *ctx, int slot);
/* The same emission on the device, so that a filling fountain needs no host transfer per step: templates = the
 * records ONE step appends (position, velocity; n_templates = group * nozzles, the reference has one nozzle of
 * group = 7), max_count = m_maxParticlesCount.  With an emitter set, every step of sph_step() starts with
 * sph_emit(): template group g is appended while count < max_count - group (the reference's test, :191), ids
 * continue the running count.  n_templates = 0 removes the emitter.  A caller that mirrors the particles on the
 * host applies the same rule there (CBaseParticleSimulator::generateParticles). */
int sph_set_emitter(sph_context *ctx, const sph_particle *templates, uint32_t n_templates, uint32_t group, uint32_t max_count);
int sph_emit(sph_context *ctx, uint32_t *n_emitted); /* one emission now; n_emitted may be NULL */
/* Read back on demand into aos[id] (the host mirror is indexed by id); capacity in records. */
int sph_download_particles(sph_context *ctx, sph_particle *aos, uint32_t capacity, uint32_t *n_out);
/* Viewer bridge (≙ the read-back + updatePosition/updateVelocity loop, src/CGPUBaseParticleSimulator.cpp:84-91, which
 * was the reference's dominant cost): the state is snapshotted into device staging on the compute stream (about 15 us
 * per million particles) and copied to aos[id] on a separate copy stream while later steps run.  aos should be
 * page-locked (sph_pin_host_buffer); its contents are defined once sph_download_wait returns.  One download can be
 * in flight - starting another first waits (on the device, not the host) for the previous copy.  Needs the particle
 * count to fit the staging buffer (4 Mi records), else SPH_ERR_ARGUMENT: use the blocking call. */
int sph_download_particles_async(sph_context *ctx, sph_particle *aos, uint32_t capacity);
int sph_download_wait(sph_context *ctx, uint32_t *n_out); /* no-op when nothing is in flight; n_out may be NULL */
int sph_particle_count(const sph_context *ctx, uint32_t *n_out);
/* ≙ CGPUBaseParticleSimulator::setGravityVector, src/CGPUBaseParticleSimulator.cpp:11-15 */
int sph_set_gravity(sph_context *ctx, float gx, float gy, float gz);
/* Collision mesh for CCollisionGeometry::inverseBounce (src/CCollisionGeometry.cpp:97-115; sFace,
 * include/CCollisionGeometry.h:46-77).  The reference prepared this per-face bounce for "collisions with a general
 * object" but its step never calls it; with n_faces > 0 the force phase adds inverseBounce(position, velocity)
 * after the bounding-box term, n_faces = 0 restores the shipped behaviour.  Every vertex of every face acts as a
 * plane with the face's normal (negated and normalised like QVector3D::normalize), spring 5000.0, damping -0.9,
 * skin 0.01 — the literals of the reference. */
typedef struct sph_face {
    float normal[3];
    float v0[3], v1[3], v2[3];
} sph_face; /* 48 bytes */
int sph_set_collision_faces(sph_context *ctx, const sph_face *faces, uint32_t n_faces);
/* Page-lock the caller's host mirror once (m_clParticles is reserved up front, src/CBaseParticleSimulator.cpp:42)
 * so per-step uploads/downloads run at full PCIe speed; unpinned again by sph_destroy. */
int sph_pin_host_buffer(sph_context *ctx, void *ptr, size_t bytes);

/* ---- the five phases (≙ the enqueueKernel call sites); *ms (may be NULL) receives the device time
 * in milliseconds (≙ CLWrapper::getEventDuration, include/CLWrapper.h:50).  With ms == NULL the call
 * is asynchronous on the context's stream. ---- */
int sph_update_grid(sph_context *ctx, double *ms);       /* keys -> counting sort -> SoA reorder; ≙ src/CGPUParticleSimulator.cpp:57-139 */
int sph_density_pressure(sph_context *ctx, double *ms);  /* ≙ src/CGPUParticleSimulator.cpp:141-155 */
int sph_forces(sph_context *ctx, double *ms);            /* pressure + viscosity + gravity; ≙ src/CGPUParticleSimulator.cpp:157-173 */
int sph_collisions(sph_context *ctx, double *ms);        /* no-op: walls are fused into integrate (CPU path: src/CCPUParticleSimulator.cpp:205-209) */
int sph_integrate(sph_context *ctx, double *ms);         /* wall penalty force + integration fused; ≙ src/CGPUBaseParticleSimulator.cpp:59-97 */
/* n_steps full steps back to back on the device (CUDA graph), no host transfer; *ms = total device time. */
int sph_step(sph_context *ctx, int n_steps, double *ms);
/* The same n_steps with CUDA events between the kernel groups: phase_ms[0..2] = grid build, density pass, forces +
 * walls + integration (sums over the steps), phase_ms[3] = sum of the whole-step spans.  Direct launches, no graph. */
int sph_step_profiled(sph_context *ctx, int n_steps, double *phase_ms4);
int sph_synchronize(sph_context *ctx);

/* ---- validation taps (all arrays indexed by particle id unless stated) ---- */
int sph_download_keys(sph_context *ctx, int32_t *keys);                /* cell id of each particle as of the last update_grid */
int sph_download_permutation(sph_context *ctx, uint32_t *sorted_ids);  /* ids in canonical (cell_id, id) order */
int sph_download_cell_start(sph_context *ctx, int32_t *cell_start);    /* cells+1 entries; slab mode: local grid, and only the
                                                                          layers in use (owned + ghost + 2) are current */
int sph_download_density_pressure_accel(sph_context *ctx, float *density, float *pressure, float *accel3);
/* neighbour sets of the grid walk (r2 <= h2, self included): counts[n]; lists (may be NULL) receives the
 * concatenated ascending id lists, particle 0 first; *total = sum(counts) */
int sph_download_neighbours(sph_context *ctx, int32_t *counts, int32_t *lists, uint64_t lists_capacity, uint64_t *total);

/* The same sets DECODED FROM THE PRODUCTION HIT WORDS: what the density pass stored and the force pass consumes
 * (bit words per particle plus the overflow list), not a re-evaluation of the predicate.  Same output convention.
 * SPH_ERR_STATE when the bitmask passes are not in use (option neighbour_variant 0, or a grid narrower than 4 cells). */
int sph_download_mask_neighbours(sph_context *ctx, int32_t *counts, int32_t *lists, uint64_t lists_capacity, uint64_t *total);

/* ---- all-pairs variant (CGPUBruteParticleSimulator semantics, resources/kernels/sph_brute.cl) ---- */
int sph_brute_density_pressure(sph_context *ctx, double *ms);
int sph_brute_forces(sph_context *ctx, double *ms);
int sph_brute_neighbour_counts(sph_context *ctx, int32_t *counts);

/* ---- rollout statistics on the device: out[0]=KE, out[1..3]=COM, out[4]=max(y)+b/2, out[5]=mean speed ---- */
int sph_stats(sph_context *ctx, double *out6);

/* q-th order statistic of the fill height y + b/2 over all particles: the element of rank floor(q (n - 1)), exact to
 * box_y / 4096^2 (two histogram passes on the device).  q = 0.95 is the robust fill height of SURVEY.md §8c. */
int sph_fill_height_percentile(sph_context *ctx, double q, double *height);

/* ---- tuning / introspection ---- */
int sph_set_option(sph_context *ctx, const char *name, int value);
/* "kernel_launches", "graph_launches", "steps", "overflow_particles", "slab_far_movers", "neighbour_pairs" (sum of the
 * neighbour counts of the last density pass, self included) */
int sph_get_counter(const sph_context *ctx, const char *name, uint64_t *value);

/* ---- slab mode (multi-GPU extension; the reference is single-device): 1-D decomposition along z, the slowest
 * cell axis of CGrid::at (include/CGrid.h:27).  One process per GPU; rank 0 creates the NCCL id and the launcher
 * broadcasts it (torch.distributed / MPI / a file).  cfg describes the WHOLE tank (box, grid_res) plus rank, world,
 * nccl_id; max_particles is this rank's capacity (owned + ghost particles).  Every step exchanges, per face, the
 * two boundary layers (ghosts) and the particles that crossed the face (migration) with ncclSend/ncclRecv. ---- */
int sph_comm_unique_id(uint8_t out[128]);
/* Loop-back transport: an id for `world` slab contexts that all live in THIS process (one device or several), to be
 * passed as nccl_id.  Each rank must then be stepped from its own host thread (the exchange blocks until both
 * neighbours take part, like NCCL ranks do); counts travel through host memory, payloads as device-to-device
 * copies ordered by CUDA events.  Everything above the transport is shared with the NCCL path — it exists so that
 * k slabs == 1 GPU can be verified bit for bit on a one-GPU box. */
int sph_comm_local_id(int32_t world, uint8_t out[128]);
int sph_slab_plan(int32_t rz, int32_t world, int32_t rank, int32_t *z0, int32_t *z1); /* owned layers [z0, z1) */
int sph_slab_create(const sph_config *cfg, sph_context **out);
/* Load balance: equal layer counts are not equal work (the neighbour count varies along the tank) and every step runs at
 * the pace of the slowest slab, so the faces between slabs MOVE.  With every exchange a rank also sends the record
 * {busy microseconds of its last complete step, z0, z1, may-grow}; both ranks at a face evaluate this pure function on
 * the same two records and move the face by its result: -1 = the lower rank gives its top layer to the upper rank,
 * +1 = the other way, 0 = stay (0.4 % hysteresis, at least 10 layers per slab, at most shift_max layers from the initial
 * plan, only into a rank that may grow).  mode 1 = by load, 2 = deterministic test pattern, 0 = never.  A moved face
 * needs no extra message: the giving rank packs the layer as "beyond the face" and keeps its copy as a ghost layer.
 * Results do not depend on where the faces are.  Option "slab_rebalance" selects the mode;
 * counters "slab_face_moves", "slab_load_us". */
int sph_slab_face_shift(const int32_t lower4[4], const int32_t upper4[4], int32_t shift, int32_t shift_max, int32_t mode,
                        uint64_t exchange, int32_t face);
/* out8 = rank, world, z0, z1, first local layer, local layers, owned particles, mean particles sent per step */
int sph_slab_info(const sph_context *ctx, int32_t out8[8]);
/* Owned particles in canonical order (not indexed by id; ids are in the records).  Works without slab mode too. */
int sph_download_owned(sph_context *ctx, sph_particle *aos, uint32_t capacity, uint32_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* SPH_CUDA_H */
