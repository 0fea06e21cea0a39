/*
 * sph_oracle.cpp — CPU ORACLE: restatement of the reference's CCPUParticleSimulator.
 *
 * TEST INFRASTRUCTURE ONLY (see sph_oracle.h).  PARITY PINNED: this file must match, bit for bit, the
 * reference's own unmodified sources compiled into oracle/_ref (tests/test_ref_pins_oracle.py) and the
 * vectors that build produced (tests/golden/ref_*.npz).
 *
 * Build: g++ -O2 -ffp-contract=off (no -ffast-math, no -march FMA) so that every fp32/fp64
 * operation rounds exactly as the reference's x86-64 build does.
 *
 * Arithmetic conventions restated from Qt 5's QVector3D (not under /root/reference; QtGui,
 * version unpinned by CMakeLists.txt:60): components are float; +,- componentwise fp32;
 * vec*float, float*vec componentwise fp32; vec/float componentwise true fp32 division;
 * dotProduct = (x1*x2 + y1*y2) + z1*z2 in fp32; lengthSquared = (x*x + y*y) + z*z in fp32;
 * a double operand is narrowed to float first because only float overloads exist.
 *
 * All file:line citations are relative to /root/reference.
 */
#include "sph_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

namespace {

// ---- QVector3D semantics -------------------------------------------------------------------
struct V3 {
    float x, y, z;
    V3() : x(0.0f), y(0.0f), z(0.0f) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(const V3 &a, float f) { return V3(a.x * f, a.y * f, a.z * f); }
inline V3 operator*(float f, const V3 &a) { return V3(a.x * f, a.y * f, a.z * f); }
inline V3 operator/(const V3 &a, float d) { return V3(a.x / d, a.y / d, a.z / d); }
inline V3 &operator+=(V3 &a, const V3 &b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline V3 &operator*=(V3 &a, float f) { a.x *= f; a.y *= f; a.z *= f; return a; }
inline float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length_squared(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

// ---- constants: include/CParticle.h:80-84, include/CCollisionGeometry.h:20-21 -------------
constexpr float kH = 0.0457f;
constexpr float kViscosity = 3.5f;
constexpr float kMass = 0.02f;
constexpr float kGasStiffness = 3.0f;
constexpr float kRestDensity = 998.29f;
constexpr double kWallK = 10000.0;
constexpr double kWallDamping = -0.9;
constexpr float kGravityAcceleration = -9.80665f;  // include/CBaseParticleSimulator.h:19

struct Wall {  // include/CCollisionGeometry.h:23-30
    V3 normal, position;
};

struct Face {  // include/CCollisionGeometry.h:46-77 (sFace: three vertices and one normal)
    V3 normal, v[3];
};

// QVector3D::normalize() of Qt 5 (qvector3d.cpp; Qt is a third-party dependency, version unpinned): squared
// length in fp64; returns unchanged if qFuzzyIsNull(len - 1.0f) or qFuzzyIsNull(len) (|d| <= 1e-12); else every
// component is divided in fp64 by sqrt(len) and narrowed.
inline void qt_normalize(V3 &a) {
    double len = (double)a.x * (double)a.x + (double)a.y * (double)a.y + (double)a.z * (double)a.z;
    if (std::fabs(len - 1.0f) <= 0.000000000001 || std::fabs(len) <= 0.000000000001) return;
    len = std::sqrt(len);
    a.x = (float)((double)a.x / len);
    a.y = (float)((double)a.y / len);
    a.z = (float)((double)a.z / len);
}

double now_ms() {
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

}  // namespace

struct OracleSim {
    // CBaseParticleSimulator state (include/CBaseParticleSimulator.h:80-90)
    V3 box;
    V3 gravity;
    float dt;
    int scenario;
    int res[3];
    int64_t max_count = 0;
    // per particle (index == id, as in m_clParticles / CParticle)
    std::vector<V3> pos, vel, acc, acc_sph, acc_wall, acc_mesh;
    std::vector<float> density, pressure, acc_scale;
    // CGrid: one vector of particle ids per cell (include/CGrid.h:24-28, src/CGrid.cpp:17-18)
    std::vector<std::vector<int32_t>> cells;
    Wall walls[6];
    std::vector<Face> faces;  // optional collision mesh (row f4); empty = the shipped behaviour
    int threads = 1;  // 1 = the reference's behaviour (single-threaded); > 1 only for the labelled all-cores baseline
    // kernel coefficients, function-local statics in the reference (src/CCPUParticleSimulator.cpp:11,19,26)
    double poly6, spiky, visc, h_squared;

    std::vector<int32_t> &at(int x, int y, int z) { return cells[(size_t)x + (size_t)y * res[0] + (size_t)z * res[0] * res[1]]; }
    const std::vector<int32_t> &at(int x, int y, int z) const { return cells[(size_t)x + (size_t)y * res[0] + (size_t)z * res[0] * res[1]]; }
    int64_t n() const { return (int64_t)pos.size(); }

    void cell_of(const V3 &p, int c[3]) const {
        // src/CCPUParticleSimulator.cpp:46-70 — fp64 math on fp32 inputs, then clamp
        double q[3] = {((double)p.x + (double)box.x / 2.0) / (double)kH,
                       ((double)p.y + (double)box.y / 2.0) / (double)kH,
                       ((double)p.z + (double)box.z / 2.0) / (double)kH};
        for (int a = 0; a < 3; ++a) {
            int v = (int)std::floor(q[a]);
            if (v < 0) v = 0;
            else if (v >= res[a]) v = res[a] - 1;
            c[a] = v;
        }
    }

    // src/CCollisionGeometry.cpp:117-133
    V3 wall_bounce(const V3 &p, const V3 &v, double *scale) const {
        V3 a(0, 0, 0);
        for (const Wall &w : walls) {
            V3 inv = w.normal * (-1.0f);
            double d = (double)dot(w.position - p, inv) + 0.01;
            if (d > 0.0) {
                V3 spring = ((float)kWallK * inv) * (float)d;
                a += spring;
                V3 damp = (float)(kWallDamping * (double)dot(v, inv)) * inv;
                a += damp;
                if (scale) *scale += std::sqrt((double)length_squared(spring)) + std::sqrt((double)length_squared(damp));
            }
        }
        return a;
    }

    // src/CCollisionGeometry.cpp:97-115 — the per-face bounce the reference prepared for "collisions with a general
    // object" but never calls: every vertex of every face is a plane with the face normal, spring 5000.0.
    V3 mesh_bounce(const V3 &p, const V3 &v, double *scale) const {
        V3 a(0, 0, 0);
        for (const Face &f : faces) {
            for (const V3 &vertex : f.v) {
                V3 inv = f.normal * (-1.0f);
                qt_normalize(inv);
                double d = (double)dot(vertex - p, inv) + 0.01;
                if (d > 0.0) {
                    V3 spring = ((float)5000.0 * inv) * (float)d;
                    a += spring;
                    V3 damp = (float)(-0.9 * (double)dot(v, inv)) * inv;
                    a += damp;
                    if (scale) *scale += std::sqrt((double)length_squared(spring)) + std::sqrt((double)length_squared(damp));
                }
            }
        }
        return a;
    }
};

// The functions below get C linkage from their declarations in sph_oracle.h.

OracleSim *oracle_create(float bx, float by, float bz, int scenario) {
    OracleSim *s = new OracleSim();
    s->box = V3(bx, by, bz);
    s->gravity = V3(0, kGravityAcceleration, 0);
    s->dt = 0.01f;  // src/CBaseParticleSimulator.cpp:7
    s->scenario = scenario;
    // src/CBaseParticleSimulator.cpp:27-31 — float division, ceil, int
    s->res[0] = (int)std::ceil(bx / kH);
    s->res[1] = (int)std::ceil(by / kH);
    s->res[2] = (int)std::ceil(bz / kH);
    s->cells.resize((size_t)s->res[0] * s->res[1] * s->res[2]);
    // coefficients as the CPU path computes them (fp64 statics)
    s->poly6 = 315.0 / (64.0 * M_PI * std::pow((double)kH, 9));
    s->spiky = -45.0 / (M_PI * std::pow((double)kH, 6));
    s->visc = 45.0 / (M_PI * std::pow((double)kH, 6));
    s->h_squared = (double)(kH * kH);  // float product, widened (src/CCPUParticleSimulator.cpp:12)
    // walls of the cuboid's bounding box (include/CCollisionGeometry.h:79-120): +-extent/2 in fp32
    V3 mn(-(bx / 2.0f), -(by / 2.0f), -(bz / 2.0f)), mx(bx / 2.0f, by / 2.0f, bz / 2.0f);
    s->walls[0] = {V3(-1, 0, 0), V3(mn.x, 0, 0)};  // left
    s->walls[1] = {V3(0, -1, 0), V3(0, mn.y, 0)};  // bottom
    s->walls[2] = {V3(0, 0, -1), V3(0, 0, mn.z)};  // back
    s->walls[3] = {V3(1, 0, 0), V3(mx.x, 0, 0)};   // right
    s->walls[4] = {V3(0, 1, 0), V3(0, mx.y, 0)};   // top
    s->walls[5] = {V3(0, 0, 1), V3(0, 0, mx.z)};   // front
    return s;
}

void oracle_destroy(OracleSim *s) { delete s; }

void oracle_add_particle(OracleSim *s, float x, float y, float z, float vx, float vy, float vz) {
    // src/CBaseParticleSimulator.cpp:67-74 — every new particle starts in cell (0,0,0)
    int32_t id = (int32_t)s->pos.size();
    s->pos.emplace_back(x, y, z);
    s->vel.emplace_back(vx, vy, vz);
    s->acc.emplace_back(0.0f, 0.0f, 0.0f);
    s->acc_sph.emplace_back(0.0f, 0.0f, 0.0f);
    s->acc_wall.emplace_back(0.0f, 0.0f, 0.0f);
    s->acc_mesh.emplace_back(0.0f, 0.0f, 0.0f);
    s->acc_scale.push_back(0.0f);
    s->density.push_back(0.0f);
    s->pressure.push_back(0.0f);
    s->at(0, 0, 0).push_back(id);
}

static unsigned calculated_count(const OracleSim *s) {
    // src/CBaseParticleSimulator.cpp:40-41
    double halfParticle = kH / 2.0f;
    return (unsigned)(std::ceil(s->box.z / halfParticle) * std::ceil(s->box.y / halfParticle) *
                      std::ceil(s->box.x / 4 / halfParticle));
}

void oracle_setup_scene(OracleSim *s) {
    // src/CBaseParticleSimulator.cpp:38-65
    double halfParticle = kH / 2.0f;
    unsigned calculatedCount = calculated_count(s);
    if (s->scenario == ORACLE_DAM_BREAK) {
        V3 offset = (-s->box) / 2.0f;
        s->pos.reserve(calculatedCount);
        for (float y = 0; y < s->box.y; y += halfParticle)
            for (float x = 0; x < s->box.x / 4.0; x += halfParticle)
                for (float z = 0; z < s->box.z; z += halfParticle)
                    oracle_add_particle(s, x + offset.x, y + offset.y, z + offset.z, 0, 0, 0);
        s->max_count = s->n();
    } else {
        s->max_count = calculatedCount;
    }
}

void oracle_set_state(OracleSim *s, int64_t n, const float *pos, const float *vel) {
    for (auto &c : s->cells) c.clear();
    s->pos.clear(); s->vel.clear(); s->acc.clear(); s->acc_sph.clear(); s->acc_wall.clear(); s->acc_mesh.clear();
    s->acc_scale.clear(); s->density.clear(); s->pressure.clear();
    for (int64_t i = 0; i < n; ++i)
        oracle_add_particle(s, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    if (s->max_count < n) s->max_count = n;
}

int oracle_overwrite_state(OracleSim *s, int64_t n, const float *pos, const float *vel) {
    if (n != s->n()) return -1;
    for (int64_t i = 0; i < n; ++i) {
        s->pos[i] = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        s->vel[i] = V3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    }
    return 0;
}

void oracle_set_gravity(OracleSim *s, float gx, float gy, float gz) { s->gravity = V3(gx, gy, gz); }

void oracle_set_faces(OracleSim *s, int n, const float *f) {
    s->faces.clear();
    std::fill(s->acc_mesh.begin(), s->acc_mesh.end(), V3());
    for (int k = 0; k < n; ++k, f += 12) {
        Face face;
        face.normal = V3(f[0], f[1], f[2]);
        for (int v = 0; v < 3; ++v) face.v[v] = V3(f[3 + 3 * v], f[4 + 3 * v], f[5 + 3 * v]);
        s->faces.push_back(face);
    }
}

int oracle_set_threads(OracleSim *s, int threads) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    s->threads = threads < 1 ? 1 : threads;
    return s->threads;
}

int oracle_generate_particles(OracleSim *s) {
    // src/CBaseParticleSimulator.cpp:187-210
    if (s->scenario != ORACLE_FOUNTAIN) return 0;
    const int particlesPerIteration = 7;
    if ((unsigned)(int32_t)s->n() >= ((unsigned)s->max_count - (unsigned)particlesPerIteration)) return 0;
    const float halfParticle = kH / 2.0f;
    V3 offset = (-s->box) / 2.0f;
    float vy = s->box.y * 3.2f;
    oracle_add_particle(s, 0, offset.y, 0, 0.0f, vy, 0.0f);
    oracle_add_particle(s, -halfParticle, offset.y, 0, 0.0f, vy, 0.0f);
    oracle_add_particle(s, halfParticle, offset.y, 0, 0.0f, vy, 0.0f);
    oracle_add_particle(s, -kH / 4, offset.y, -halfParticle, 0.0f, vy, 0.0f);
    oracle_add_particle(s, kH / 4, offset.y, -halfParticle, 0.0f, vy, 0.0f);
    oracle_add_particle(s, -kH / 4, offset.y, halfParticle, 0.0f, vy, 0.0f);
    oracle_add_particle(s, kH / 4, offset.y, halfParticle, 0.0f, vy, 0.0f);
    return particlesPerIteration;
}

double oracle_update_grid(OracleSim *s) {
    // src/CCPUParticleSimulator.cpp:32-91 — traversal x,y,z; swap-with-last + pop; redo the index
    double t0 = now_ms();
    for (int x = 0; x < s->res[0]; x++)
        for (int y = 0; y < s->res[1]; y++)
            for (int z = 0; z < s->res[2]; z++) {
                std::vector<int32_t> &particles = s->at(x, y, z);
                for (size_t p = 0; p < particles.size(); p++) {
                    int32_t id = particles[p];
                    int c[3];
                    s->cell_of(s->pos[id], c);
                    if (x != c[0] || y != c[1] || z != c[2]) {
                        s->at(c[0], c[1], c[2]).push_back(id);
                        particles[p] = particles.back();
                        particles.pop_back();
                        p--;
                    }
                }
            }
    return now_ms() - t0;
}

// Visit the (up to) 27 cells around (x,y,z) in the reference's order and call f(neighbour id).
template <typename F>
static inline void for_neighbours(const OracleSim *s, int x, int y, int z, F &&f) {
    // src/CCPUParticleSimulator.cpp:107-131 (continue below 0, break at >= res)
    for (int ox = -1; ox <= 1; ox++) {
        if (x + ox < 0) continue;
        if (x + ox >= s->res[0]) break;
        for (int oy = -1; oy <= 1; oy++) {
            if (y + oy < 0) continue;
            if (y + oy >= s->res[1]) break;
            for (int oz = -1; oz <= 1; oz++) {
                if (z + oz < 0) continue;
                if (z + oz >= s->res[2]) break;
                for (int32_t j : s->at(x + ox, y + oy, z + oz)) f(j);
            }
        }
    }
}

// Cell traversal of the density and force loops: x, y, z nested like the reference
// (src/CCPUParticleSimulator.cpp:98-100).  With threads > 1 (NOT reference behaviour, labelled all-cores baseline)
// the (x, y) columns are handed out to std::threads; every particle's sum only reads shared state and runs in the
// same order as in the serial loop, so the results are bit-identical.
template <typename F>
static inline void for_cells(OracleSim *s, F &&f) {
    const int nx = s->res[0], ny = s->res[1], nz = s->res[2];
    if (s->threads <= 1) {
        for (int x = 0; x < nx; x++)
            for (int y = 0; y < ny; y++)
                for (int z = 0; z < nz; z++) f(x, y, z);
        return;
    }
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (int c = next.fetch_add(1); c < nx * ny; c = next.fetch_add(1))
            for (int z = 0; z < nz; z++) f(c / ny, c % ny, z);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < s->threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

static inline void density_add(OracleSim *s, int32_t i, int32_t j) {
    // src/CCPUParticleSimulator.cpp:122-127 and Wpoly6 :9-15
    V3 distance = s->pos[i] - s->pos[j];
    double radiusSquared = length_squared(distance);
    if (radiusSquared <= kH * kH) {
        double w = s->poly6 * std::pow(s->h_squared - radiusSquared, 3);
        s->density[i] = (float)((double)s->density[i] + w);  // float += double
    }
}

static inline void density_finish(OracleSim *s, int32_t i) {
    // src/CCPUParticleSimulator.cpp:133-134
    s->density[i] *= kMass;
    s->pressure[i] = kGasStiffness * (s->density[i] - kRestDensity);
}

double oracle_update_density_pressure(OracleSim *s) {
    // src/CCPUParticleSimulator.cpp:93-141
    double t0 = now_ms();
    for_cells(s, [&](int x, int y, int z) {
        for (int32_t i : s->at(x, y, z)) {
            s->density[i] = 0.0f;
            for_neighbours(s, x, y, z, [&](int32_t j) { density_add(s, i, j); });
            density_finish(s, i);
        }
    });
    return now_ms() - t0;
}

struct ForceAcc {
    V3 f_pressure, f_viscosity;
    double mag_p = 0.0, mag_v = 0.0;  // sums of term magnitudes (tolerance scale only)
};

static inline void force_add(const OracleSim *s, int32_t i, int32_t j, ForceAcc &fa) {
    // src/CCPUParticleSimulator.cpp:174-186; WspikyGradient :17-22; WviscosityLaplacian :24-30
    V3 distance = s->pos[i] - s->pos[j];
    double radiusSquared = length_squared(distance);
    if (radiusSquared <= kH * kH && i != j) {
        double radius = std::sqrt(radiusSquared);
        V3 spikyGradient = ((float)(s->spiky * std::pow((double)kH - radius, 2)) * distance) / (float)radius;
        double viscosityLaplacian = s->visc * ((double)kH - radius);
        double scalar = (double)s->pressure[i] / std::pow((double)s->density[i], 2) +
                        ((double)s->pressure[j] / std::pow((double)s->density[j], 2));
        V3 tp = (float)scalar * spikyGradient;
        fa.f_pressure += tp;
        V3 tv = ((s->vel[j] - s->vel[i]) * (float)viscosityLaplacian) / s->density[j];
        fa.f_viscosity += tv;
        fa.mag_p += std::sqrt((double)length_squared(tp));
        fa.mag_v += std::sqrt((double)length_squared(tv));
    }
}

static inline void force_finish(OracleSim *s, int32_t i, ForceAcc &fa) {
    // src/CCPUParticleSimulator.cpp:153,191-198
    float rho = s->density[i];
    V3 f_gravity = s->gravity * rho;
    fa.f_pressure *= -kMass * rho;
    fa.f_viscosity *= kViscosity * kMass;
    V3 a = (fa.f_pressure + fa.f_viscosity + f_gravity) / rho;
    s->acc_sph[i] = a;
    double scale = ((double)kMass * rho * fa.mag_p + (double)(kViscosity * kMass) * fa.mag_v +
                    std::sqrt((double)length_squared(f_gravity))) / (double)rho;
    const V3 wall = s->wall_bounce(s->pos[i], s->vel[i], &scale);
    s->acc_wall[i] = wall;
    a += wall;
    // extension point (row f4): the reference never calls inverseBounce; with a mesh set, its result is added the
    // same way the bounding-box term is (`acceleration() += ...`, src/CCPUParticleSimulator.cpp:196)
    if (!s->faces.empty()) {
        const V3 mesh = s->mesh_bounce(s->pos[i], s->vel[i], &scale);
        s->acc_mesh[i] = mesh;
        a += mesh;
    }
    s->acc[i] = a;
    s->acc_scale[i] = (float)scale;
}

double oracle_update_forces(OracleSim *s) {
    // src/CCPUParticleSimulator.cpp:143-203
    double t0 = now_ms();
    for_cells(s, [&](int x, int y, int z) {
        for (int32_t i : s->at(x, y, z)) {
            ForceAcc fa;
            for_neighbours(s, x, y, z, [&](int32_t j) { force_add(s, i, j, fa); });
            force_finish(s, i, fa);
        }
    });
    return now_ms() - t0;
}

double oracle_update_collisions(OracleSim *) { return 0.0; }  // src/CCPUParticleSimulator.cpp:205-209

double oracle_integrate(OracleSim *s) {
    // src/CCPUParticleSimulator.cpp:211-229 — cell order is irrelevant (per-particle update)
    double t0 = now_ms();
    const float dt = s->dt;
    for (auto &cell : s->cells)
        for (int32_t i : cell) {
            V3 newPosition = s->pos[i] + (s->vel[i] * dt) + s->acc[i] * dt * dt;
            V3 newVelocity = (newPosition - s->pos[i]) / dt;
            s->pos[i] = newPosition;
            s->vel[i] = newVelocity;
        }
    return now_ms() - t0;
}

void oracle_step(OracleSim *s, int n_steps, double *phase_ms) {
    // src/CBaseParticleSimulator.cpp:116-144
    for (int k = 0; k < n_steps; ++k) {
        oracle_generate_particles(s);
        double g = oracle_update_grid(s);
        double d = oracle_update_density_pressure(s);
        double f = oracle_update_forces(s);
        double c = oracle_update_collisions(s);
        double in = oracle_integrate(s);
        if (phase_ms) { phase_ms[0] += g; phase_ms[1] += d; phase_ms[2] += f; phase_ms[3] += c; phase_ms[4] += in; }
    }
}

void oracle_brute_density_pressure(OracleSim *s) {
    // resources/kernels/sph_brute.cl:9-31 with the CPU path's arithmetic; neighbours in index order
    const int32_t n = (int32_t)s->n();
    for (int32_t i = 0; i < n; ++i) {
        s->density[i] = 0.0f;
        for (int32_t j = 0; j < n; ++j) density_add(s, i, j);
        density_finish(s, i);
    }
}

void oracle_brute_forces(OracleSim *s) {
    // resources/kernels/sph_brute.cl:33-57 (+ the CPU path's wall term)
    const int32_t n = (int32_t)s->n();
    for (int32_t i = 0; i < n; ++i) {
        ForceAcc fa;
        for (int32_t j = 0; j < n; ++j) force_add(s, i, j, fa);
        force_finish(s, i, fa);
    }
}

int64_t oracle_count(const OracleSim *s) { return s->n(); }
int64_t oracle_max_count(const OracleSim *s) { return s->max_count; }
void oracle_grid_res(const OracleSim *s, int *r) { r[0] = s->res[0]; r[1] = s->res[1]; r[2] = s->res[2]; }
void oracle_constants(const OracleSim *s, double *poly6, double *spiky, double *visc, float *h2) {
    *poly6 = s->poly6; *spiky = s->spiky; *visc = s->visc; *h2 = kH * kH;
}

static void copy3(const std::vector<V3> &v, float *out) {
    for (size_t i = 0; i < v.size(); ++i) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
}
void oracle_get_pos(const OracleSim *s, float *o) { copy3(s->pos, o); }
void oracle_get_vel(const OracleSim *s, float *o) { copy3(s->vel, o); }
void oracle_get_acc(const OracleSim *s, float *o) { copy3(s->acc, o); }
void oracle_get_acc_sph(const OracleSim *s, float *o) { copy3(s->acc_sph, o); }
void oracle_get_acc_wall(const OracleSim *s, float *o) { copy3(s->acc_wall, o); }
void oracle_get_acc_mesh(const OracleSim *s, float *o) { copy3(s->acc_mesh, o); }
void oracle_get_acc_scale(const OracleSim *s, float *o) { std::memcpy(o, s->acc_scale.data(), s->acc_scale.size() * sizeof(float)); }
void oracle_get_density(const OracleSim *s, float *o) { std::memcpy(o, s->density.data(), s->density.size() * sizeof(float)); }
void oracle_get_pressure(const OracleSim *s, float *o) { std::memcpy(o, s->pressure.data(), s->pressure.size() * sizeof(float)); }

void oracle_get_keys(const OracleSim *s, int32_t *out) {
    for (int64_t i = 0; i < s->n(); ++i) {
        int c[3];
        s->cell_of(s->pos[i], c);
        out[i] = c[0] + c[1] * s->res[0] + c[2] * s->res[0] * s->res[1];
    }
}

void oracle_get_cells(const OracleSim *s, int32_t *cell_start, int32_t *ids) {
    int32_t run = 0;
    for (size_t c = 0; c < s->cells.size(); ++c) {
        cell_start[c] = run;
        std::vector<int32_t> m(s->cells[c]);
        std::sort(m.begin(), m.end());
        for (int32_t id : m) ids[run++] = id;
    }
    cell_start[s->cells.size()] = run;
}

void oracle_get_cells_raw(const OracleSim *s, int32_t *cell_start, int32_t *ids) {
    int32_t run = 0;
    for (size_t c = 0; c < s->cells.size(); ++c) {
        cell_start[c] = run;
        for (int32_t id : s->cells[c]) ids[run++] = id;
    }
    cell_start[s->cells.size()] = run;
}

int64_t oracle_get_neighbours(const OracleSim *s, int32_t *counts, int32_t *lists) {
    std::vector<std::vector<int32_t>> nb((size_t)s->n());
    for (int x = 0; x < s->res[0]; x++)
        for (int y = 0; y < s->res[1]; y++)
            for (int z = 0; z < s->res[2]; z++)
                for (int32_t i : s->at(x, y, z))
                    for_neighbours(s, x, y, z, [&](int32_t j) {
                        V3 d = s->pos[i] - s->pos[j];
                        double r2 = length_squared(d);
                        if (r2 <= kH * kH) nb[i].push_back(j);
                    });
    int64_t total = 0;
    for (int64_t i = 0; i < s->n(); ++i) {
        std::sort(nb[i].begin(), nb[i].end());
        counts[i] = (int32_t)nb[i].size();
        if (lists) std::memcpy(lists + total, nb[i].data(), nb[i].size() * sizeof(int32_t));
        total += (int64_t)nb[i].size();
    }
    return total;
}

void oracle_stats(const OracleSim *s, double *out) {
    // definitions of SURVEY.md §8c (the reference computes none of these)
    const int64_t n = s->n();
    double ke = 0, cx = 0, cy = 0, cz = 0, ymax = -1e300;
    std::vector<float> ys((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const V3 &v = s->vel[i], &p = s->pos[i];
        ke += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z;
        cx += p.x; cy += p.y; cz += p.z;
        if (p.y > ymax) ymax = p.y;
        ys[i] = p.y;
    }
    double p95 = 0;
    if (n > 0) {
        size_t k = (size_t)(0.95 * (double)(n - 1));
        std::nth_element(ys.begin(), ys.begin() + k, ys.end());
        p95 = ys[k];
    }
    double hb = (double)s->box.y / 2.0;
    out[0] = 0.5 * (double)kMass * ke;
    out[1] = n ? cx / n : 0; out[2] = n ? cy / n : 0; out[3] = n ? cz / n : 0;
    out[4] = n ? ymax + hb : 0;
    out[5] = n ? p95 + hb : 0;
}

