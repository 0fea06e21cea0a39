// sim_capi.cpp — C facade over the C++ simulator objects so that Python (tests, bench.py) drives the
// very same CCUDAParticleSimulator a C++ caller would.  Not part of the reference-facing boundary
// (that is include/sph_cuda.h); declared in include/sph_host.h.
#include "../../include/sph_host.h"

#include <cstdio>
#include <cstring>
#include <string>

#include "ExportLogs.h"
#include "SimulatorFactory.h"

namespace {
thread_local std::string g_err;

// Scene generation needs no device: a simulator whose phases do nothing exposes setupScene() and the
// fountain emitter for host-logic tests that run without a GPU.
class CSceneOnlySimulator : public CBaseParticleSimulator {
public:
    using CBaseParticleSimulator::CBaseParticleSimulator;
    QString getSelectedDevice() override { return "none (scene generation only)"; }

protected:
    double updateGrid() override { return 0; }
    double updateDensityPressure() override { return 0; }
    double updateForces() override { return 0; }
    double updateCollisions() override { return 0; }
    double integrate() override { return 0; }
};

struct Handle {
    CScene scene;
    CBaseParticleSimulator *sim = nullptr;
    CCUDAParticleSimulator *cuda = nullptr;  // same object when it is a CUDA simulator
    std::string device_name, last_error;
    int errors = 0;
    ~Handle() { delete sim; }
};

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

const char *gmu_sim_last_error(void) { return g_err.c_str(); }

gmu_sim *gmu_sim_create(const char *type, float bx, float by, float bz, int device, int scenario) {
    Handle *h = new Handle();
    const std::string t(type ? type : "");
    int rc = guarded([&] {
        const QVector3D box(bx, by, bz);
        const auto sc = (SimulationScenario)scenario;
        if (t == "scene_only") {
            h->sim = new CSceneOnlySimulator(&h->scene, box, sc);
        } else if (t == "cuda" || t == "cuda_grid") {
            h->sim = h->cuda = static_cast<CCUDAParticleSimulator *>(createSimulator(eSimulationType::CUDAGrid, &h->scene, box, device, sc));
        } else if (t == "cuda_brute") {
            h->sim = h->cuda = static_cast<CCUDAParticleSimulator *>(createSimulator(eSimulationType::CUDABrute, &h->scene, box, device, sc));
        } else {
            throw std::invalid_argument("gmu_sim_create: unknown simulator type '" + t + "'");
        }
    });
    if (rc) {
        delete h;
        return nullptr;
    }
    // errors inside step() arrive as the errorOccured signal, like in the reference
    h->sim->onErrorOccured([h](const char *what) { h->last_error = what; ++h->errors; });
    return reinterpret_cast<gmu_sim *>(h);
}

void gmu_sim_destroy(gmu_sim *s) { delete reinterpret_cast<Handle *>(s); }

#define H(s) (reinterpret_cast<Handle *>(s))

int gmu_sim_setup_scene(gmu_sim *s) { return guarded([&] { H(s)->sim->setupScene(); }); }

int gmu_sim_step(gmu_sim *s, int n) {
    return guarded([&] {
        H(s)->errors = 0;
        for (int k = 0; k < n && !H(s)->errors; ++k) H(s)->sim->doWork();
        if (H(s)->errors) throw std::runtime_error(H(s)->last_error);
    });
}

int gmu_sim_step_many(gmu_sim *s, int n, double *device_ms) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_step_many: not a CUDA simulator");
        H(s)->cuda->stepMany(n, device_ms);
    });
}

int gmu_sim_emit(gmu_sim *s, int n_steps) {
    // run only the emission part of step() n times (scene_only simulators: phases are no-ops)
    return guarded([&] { for (int k = 0; k < n_steps; ++k) H(s)->sim->doWork(); });
}

int gmu_sim_set_owned_layers(gmu_sim *s, int z0, int z1) {
    return guarded([&] { H(s)->sim->setOwnedLayers(z0, z1); });
}

int gmu_sim_enable_slab(gmu_sim *s, int rank, int world, const unsigned char *nccl_id128) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_enable_slab: not a CUDA simulator");
        H(s)->cuda->enableSlab(rank, world, nccl_id128);
    });
}

int gmu_sim_set_mirror_mode(gmu_sim *s, int mode) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_set_mirror_mode: not a CUDA simulator");
        H(s)->cuda->setMirrorMode((CCUDAParticleSimulator::MirrorMode)mode);
    });
}

int gmu_sim_set_mirror_stride(gmu_sim *s, int stride) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_set_mirror_stride: not a CUDA simulator");
        H(s)->cuda->setMirrorStride(stride);
    });
}

int gmu_sim_sync_host(gmu_sim *s) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_sync_host: not a CUDA simulator");
        H(s)->cuda->syncHostMirror();
    });
}

int gmu_sim_set_gravity(gmu_sim *s, float gx, float gy, float gz) {
    return guarded([&] { H(s)->sim->setGravityVector(QVector3D(gx, gy, gz)); });
}

int gmu_sim_set_collision_faces(gmu_sim *s, const float *f, int n_faces) {
    return guarded([&] {
        if (!H(s)->cuda) throw std::runtime_error("gmu_sim_set_collision_faces: not a CUDA simulator");
        std::vector<sFace> faces;
        for (int k = 0; k < n_faces; ++k, f += 12) {
            const QVector3D normal(f[0], f[1], f[2]);
            faces.emplace_back(sVertex(QVector3D(f[3], f[4], f[5]), normal), sVertex(QVector3D(f[6], f[7], f[8]), normal),
                               sVertex(QVector3D(f[9], f[10], f[11]), normal));  // the face takes its first vertex's normal
        }
        H(s)->cuda->setCollisionFaces(faces);
    });
}

int gmu_sim_get_gravity(gmu_sim *s, float *out3) {
    return guarded([&] {
        const QVector3D g = H(s)->sim->getGravityVector();
        out3[0] = g.x(); out3[1] = g.y(); out3[2] = g.z();
    });
}

int gmu_sim_is_running(gmu_sim *s) { return H(s)->sim->running() ? 1 : 0; }

int gmu_sim_key(gmu_sim *s, int qt_key) { return guarded([&] { H(s)->sim->onKeyPressed((Qt::Key)qt_key); }); }

int gmu_sim_set_profiling(gmu_sim *s, int on, int stride) {
    return guarded([&] {
        H(s)->sim->setProfiling(on != 0);
        if (stride > 0) H(s)->sim->eventLoggerStride = stride;
    });
}

int gmu_sim_set_emission_multiplier(gmu_sim *s, int nozzles) {
    return guarded([&] { H(s)->sim->setEmissionMultiplier(nozzles); });
}

uint64_t gmu_sim_particle_count(gmu_sim *s) { return H(s)->sim->getParticlesCount(); }
uint64_t gmu_sim_max_particle_count(gmu_sim *s) { return H(s)->sim->getMaxParticlesCount(); }
uint64_t gmu_sim_iteration(gmu_sim *s) { return H(s)->sim->getTotalIteration(); }

int gmu_sim_wait_host(gmu_sim *s) {
    return guarded([&] {
        if (H(s)->cuda) H(s)->cuda->waitHostMirror();
    });
}

const sph_particle *gmu_sim_host_particles(gmu_sim *s) {
    if (H(s)->cuda) guarded([&] { H(s)->cuda->waitHostMirror(); });  // never hand out a half-copied snapshot
    return reinterpret_cast<const sph_particle *>(H(s)->sim->getHostParticles().data());
}

sph_context *gmu_sim_context(gmu_sim *s) { return H(s)->cuda ? H(s)->cuda->context() : nullptr; }

const char *gmu_sim_device_name(gmu_sim *s) {
    if (guarded([&] { H(s)->device_name = H(s)->sim->getSelectedDevice(); })) return "";
    return H(s)->device_name.c_str();
}

uint64_t gmu_sim_event_count(gmu_sim *s) { return H(s)->sim->events.size(); }

uint64_t gmu_sim_get_events(gmu_sim *s, double *out, uint64_t max_events) {
    // 7 doubles per event: iteration, fps, grid, density, forces, collisions, integrate
    const auto &ev = H(s)->sim->events;
    uint64_t n = ev.size() < max_events ? ev.size() : max_events;
    for (uint64_t i = 0; i < n; ++i) {
        const sProfilingEvent &e = ev[i];
        double *o = out + 7 * i;
        o[0] = (double)e.iteration; o[1] = e.fps; o[2] = e.updateGrid; o[3] = e.updateDensityPressure;
        o[4] = e.updateForces; o[5] = e.updateCollisions; o[6] = e.integrate;
    }
    return n;
}

int gmu_sim_push_event(gmu_sim *s, const double *ev7) {
    // appends one profiling record (iteration, fps, grid, density, forces, collisions, integrate): lets a caller
    // merge samples measured elsewhere into the log, and the tests pin the CSV layout with known numbers
    return guarded([&] {
        sProfilingEvent e((unsigned long)ev7[0]);
        e.fps = ev7[1]; e.updateGrid = ev7[2]; e.updateDensityPressure = ev7[3]; e.updateForces = ev7[4];
        e.updateCollisions = ev7[5]; e.integrate = ev7[6];
        H(s)->sim->events << e;
    });
}

int gmu_sim_type_from_name(const char *combo_text) {
    for (int t = 0; t <= (int)eSimulationType::CUDABrute; ++t)
        if (std::string(simulationTypeName((eSimulationType)t)) == (combo_text ? combo_text : "")) return t;
    return -1;
}

gmu_sim *gmu_sim_create_by_type(int type, float bx, float by, float bz, int device, int scenario) {
    Handle *h = new Handle();
    int rc = guarded([&] {
        h->sim = createSimulator((eSimulationType)type, &h->scene, QVector3D(bx, by, bz), device, (SimulationScenario)scenario);
        h->cuda = static_cast<CCUDAParticleSimulator *>(h->sim);  // createSimulator only returns CUDA simulators
    });
    if (rc) {
        delete h;
        return nullptr;
    }
    h->sim->onErrorOccured([h](const char *what) { h->last_error = what; ++h->errors; });
    return reinterpret_cast<gmu_sim *>(h);
}

int gmu_sim_export_logs(gmu_sim *s, const char *dir, const char *sim_name) {
    return guarded([&] { exportLogs(*H(s)->sim, dir, sim_name); });
}

}  // extern "C"
