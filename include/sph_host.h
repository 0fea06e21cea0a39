/*
 * sph_host.h — C facade of the C++ simulator objects (libsph_host.so), for non-C++ drivers
 * (tests, bench.py).  The reference-facing boundary is include/sph_cuda.h; this header only lets a
 * Python process construct and drive the same CCUDAParticleSimulator a C++ caller would, the way
 * MainWindow does (src/mainwindow.cpp:171-201,243-287).
 */
#ifndef SPH_HOST_H
#define SPH_HOST_H

#include <stdint.h>

#include "sph_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmu_sim gmu_sim;

/* type: "cuda" (uniform grid), "cuda_brute" (all pairs), "scene_only" (no device; scene generation and
 * emission only).  scenario: 0 = DAM_BREAK, 1 = FOUNTAIN (include/CBaseParticleSimulator.h:21-25). */
gmu_sim *gmu_sim_create(const char *type, float box_x, float box_y, float box_z, int device, int scenario);
void gmu_sim_destroy(gmu_sim *s);
const char *gmu_sim_last_error(void);

/* multi-GPU extension: make the simulator rank `rank` of `world` z-slabs; call before gmu_sim_setup_scene */
/* host half of the slab split: keep only the particles of z-layers [z0, z1) at scene generation (any simulator type) */
int gmu_sim_set_owned_layers(gmu_sim *s, int z0, int z1);
int gmu_sim_enable_slab(gmu_sim *s, int rank, int world, const unsigned char *nccl_id128);
int gmu_sim_setup_scene(gmu_sim *s);                               /* setupScene() */
int gmu_sim_step(gmu_sim *s, int n);                               /* n timer ticks: doWork() = step() + counters */
int gmu_sim_step_many(gmu_sim *s, int n, double *device_ms);       /* fused device steps (extension) */
int gmu_sim_emit(gmu_sim *s, int n_steps);                         /* scene_only: run the emitter n steps */
int gmu_sim_set_mirror_mode(gmu_sim *s, int mode);                 /* 0 resident, 1 download, 2 round trip per step, 3 asynchronous download */
int gmu_sim_set_mirror_stride(gmu_sim *s, int stride);              /* download mode: refresh every stride-th step */
int gmu_sim_sync_host(gmu_sim *s);                                 /* device -> host mirror, blocking, current state */
int gmu_sim_wait_host(gmu_sim *s);                                 /* mode 3: complete the read-back in flight */
int gmu_sim_set_gravity(gmu_sim *s, float gx, float gy, float gz); /* setGravityVector */
int gmu_sim_get_gravity(gmu_sim *s, float *out3);                  /* the simulator's gravity member */
int gmu_sim_is_running(gmu_sim *s);                                /* isRunning(): the timer is active (start/toggleSimulation) */
int gmu_sim_key(gmu_sim *s, int qt_key);                           /* onKeyPressed */
/* collision mesh for CCollisionGeometry::inverseBounce: n faces x 12 floats (normal, v0, v1, v2); 0 clears it */
int gmu_sim_set_collision_faces(gmu_sim *s, const float *faces12, int n_faces);
int gmu_sim_set_profiling(gmu_sim *s, int on, int stride);
int gmu_sim_set_emission_multiplier(gmu_sim *s, int nozzles);
uint64_t gmu_sim_particle_count(gmu_sim *s);
uint64_t gmu_sim_max_particle_count(gmu_sim *s);
uint64_t gmu_sim_iteration(gmu_sim *s);
const sph_particle *gmu_sim_host_particles(gmu_sim *s);            /* m_clParticles.data() */
sph_context *gmu_sim_context(gmu_sim *s);                          /* the simulator's device context (taps) */
const char *gmu_sim_device_name(gmu_sim *s);                       /* getSelectedDevice() */
uint64_t gmu_sim_event_count(gmu_sim *s);
uint64_t gmu_sim_get_events(gmu_sim *s, double *out7, uint64_t max_events);
int gmu_sim_push_event(gmu_sim *s, const double *ev7);            /* append one record (same 7 doubles) to `events` */
/* MainWindow::exportLogs (src/mainwindow.cpp:310-368), byte for byte: appends "<Scenario>_<box>.csv" and
 * "<Scenario>_<box>_detail.csv" in dir; sim_name is the simulation-type combo text ("CUDA Grid") */
int gmu_sim_export_logs(gmu_sim *s, const char *dir, const char *sim_name);
/* eSimulationType (include/mainwindow.h:60-63) with the CUDA types appended: "GPU Grid" 0, "GPU Brute Force" 1,
 * "CPU Grid" 2, "CUDA Grid" 3, "CUDA Brute Force" 4; -1 for an unknown text */
int gmu_sim_type_from_name(const char *combo_text);
/* createSimulator(type, ...) (src/mainwindow.cpp:171-201); NULL + gmu_sim_last_error() for the types not built here */
gmu_sim *gmu_sim_create_by_type(int type, float box_x, float box_y, float box_z, int device, int scenario);

#ifdef __cplusplus
}
#endif
#endif /* SPH_HOST_H */
