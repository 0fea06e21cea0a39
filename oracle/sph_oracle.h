/*
 * sph_oracle.h — C interface of the CPU ORACLE for the SPH time step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * CCPUParticleSimulator (single-threaded, mixed fp32/fp64 arithmetic).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (libsph_cuda.so, libsph_host.so) never
 * links, loads or calls anything in oracle/.
 *
 * PARITY PINNED against the reference's own code: oracle/_ref/libsph_ref.so is the reference's
 * CCPUParticleSimulator / CBaseParticleSimulator / CCollisionGeometry / CGrid / CParticle compiled UNMODIFIED
 * from /root/reference against stand-in Qt headers (oracle/qt_shim, oracle/ref_driver.cpp, Makefile target
 * `ref`).  This restatement must reproduce it BIT FOR BIT every step — positions, velocities, densities,
 * pressures, accelerations and the history-dependent order inside every grid cell — on BASELINE configs[0]
 * (16 000 particles, 100 steps), the fountain, random states, tilted gravity, the wall and mesh bounce
 * (tests/test_ref_pins_oracle.py), and against the committed vectors that library produced
 * (tests/golden/ref_*.npz, tests/test_oracle_matches_ref_golden.py; these run where /root/reference is absent).
 * The one piece that stays restated on both sides is Qt's QVector3D (QtGui is not in this image; its inline
 * semantics are written out in oracle/qt_shim/qt_shim_core.h).  The reference ships no tests or golden vectors
 * of its own; known answers derived from its source are in tests/test_oracle_known_answers.py.
 *
 * All citations are relative to /root/reference.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleSim OracleSim;

enum { ORACLE_DAM_BREAK = 0, ORACLE_FOUNTAIN = 1 };

/* ctor of CBaseParticleSimulator (src/CBaseParticleSimulator.cpp:3-36); the reference only
 * takes a cube (one float); a non-cubic box is the extension needed for the 64M tank. */
OracleSim *oracle_create(float box_x, float box_y, float box_z, int scenario);
void oracle_destroy(OracleSim *s);

/* setupScene (src/CBaseParticleSimulator.cpp:38-65): dam-break lattice or empty fountain */
void oracle_setup_scene(OracleSim *s);
/* replace the whole state: n particles, ids 0..n-1, xyz-interleaved fp32 */
void oracle_set_state(OracleSim *s, int64_t n, const float *pos, const float *vel);
/* overwrite positions / velocities of the existing particles and leave the cell vectors alone (what integrate()
 * does): the next update_grid moves them with the history the cells already have.  -1 unless n == count */
int oracle_overwrite_state(OracleSim *s, int64_t n, const float *pos, const float *vel);
/* addParticle (src/CBaseParticleSimulator.cpp:67-74) */
void oracle_add_particle(OracleSim *s, float x, float y, float z, float vx, float vy, float vz);
void oracle_set_gravity(OracleSim *s, float gx, float gy, float gz);
/* collision mesh for CCollisionGeometry::inverseBounce (src/CCollisionGeometry.cpp:97-115), which the reference
 * prepared but never calls: n faces, 12 floats each = normal xyz, v0 xyz, v1 xyz, v2 xyz.  With n > 0 the force
 * phase adds inverseBounce(position, velocity) after the bounding-box term; n = 0 restores the shipped behaviour. */
void oracle_set_faces(OracleSim *s, int n, const float *faces12);
/* NOT reference behaviour (the reference CPU path is single-threaded): run the density and force loops on `threads`
 * std::threads (<= 0: all cores) for the labelled all-cores baseline; results stay bit-identical. Returns the count used. */
int oracle_set_threads(OracleSim *s, int threads);

/* generateParticles (src/CBaseParticleSimulator.cpp:187-210); returns #particles added */
int oracle_generate_particles(OracleSim *s);

/* the five phases (src/CCPUParticleSimulator.cpp:32-229); each returns elapsed ms */
double oracle_update_grid(OracleSim *s);
double oracle_update_density_pressure(OracleSim *s);
double oracle_update_forces(OracleSim *s);
double oracle_update_collisions(OracleSim *s);
double oracle_integrate(OracleSim *s);
/* step() (src/CBaseParticleSimulator.cpp:116-144), n times; phase_ms[5] accumulates (may be NULL) */
void oracle_step(OracleSim *s, int n_steps, double *phase_ms);

/* all-pairs variant with CGPUBruteParticleSimulator semantics (resources/kernels/sph_brute.cl)
 * but the CPU path's arithmetic; fills density/pressure/acceleration like the grid phases do */
void oracle_brute_density_pressure(OracleSim *s);
void oracle_brute_forces(OracleSim *s);

/* ---- state taps (arrays indexed by particle id) ---- */
int64_t oracle_count(const OracleSim *s);
int64_t oracle_max_count(const OracleSim *s);
void oracle_grid_res(const OracleSim *s, int *res3);
void oracle_constants(const OracleSim *s, double *poly6, double *spiky, double *visc, float *h2_f32);
void oracle_get_pos(const OracleSim *s, float *out3n);
void oracle_get_vel(const OracleSim *s, float *out3n);
void oracle_get_acc(const OracleSim *s, float *out3n);       /* SPH + wall term (what updateForces leaves) */
void oracle_get_acc_sph(const OracleSim *s, float *out3n);   /* before the wall term is added */
void oracle_get_acc_wall(const OracleSim *s, float *out3n);  /* the wall term alone (inverseBoundingBoxBounce) */
void oracle_get_acc_mesh(const OracleSim *s, float *out3n);  /* the mesh term alone (inverseBounce); zero without faces */
void oracle_get_acc_scale(const OracleSim *s, float *outn);  /* sum of |terms| / rho: conditioning scale for tolerances */
void oracle_get_density(const OracleSim *s, float *outn);
void oracle_get_pressure(const OracleSim *s, float *outn);
/* cell key of every particle computed from its CURRENT position (formula of updateGrid) */
void oracle_get_keys(const OracleSim *s, int32_t *outn);
/* per-cell membership of the grid as updateGrid left it: cell_start[cells+1], ids[n] sorted by id inside each cell
 * (canonical (cell,id) permutation) */
void oracle_get_cells(const OracleSim *s, int32_t *cell_start, int32_t *ids);
/* the same membership in the order the cell vectors actually hold (push_back / swap-and-pop history,
 * src/CCPUParticleSimulator.cpp:72-83): the traversal order of the density and force sums, compared with oracle/_ref */
void oracle_get_cells_raw(const OracleSim *s, int32_t *cell_start, int32_t *ids);
/* neighbour sets from the current grid: r2 <= h2, self included; counts[n]; if lists != NULL it must hold
 * sum(counts) ids, neighbours of particle 0 first, each list sorted ascending */
int64_t oracle_get_neighbours(const OracleSim *s, int32_t *counts, int32_t *lists);

/* rollout statistics: KE = 0.5*m*sum|v|^2, COM xyz, fill height max(y)+b/2, 95th percentile */
void oracle_stats(const OracleSim *s, double *out6);

#ifdef __cplusplus
}
#endif
#endif
