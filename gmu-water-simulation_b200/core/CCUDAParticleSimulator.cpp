// CCUDAParticleSimulator.cpp — host side of the CUDA simulator: scene upload, phase calls through the
// C ABI, error propagation as in CGPUBaseParticleSimulator (src/CGPUBaseParticleSimulator.cpp:28-37).
#include "CCUDAParticleSimulator.h"

#include <algorithm>
#include <cmath>
#include <cstring>

static_assert(sizeof(CParticle::Physics) == sizeof(sph_particle), "host mirror record != ABI record");

CCUDAParticleSimulator::CCUDAParticleSimulator(CScene *scene, float boxSize, int device, SimulationScenario scenario, QObject *parent)
    : CBaseParticleSimulator(scene, boxSize, scenario, parent), m_device(device) {}

CCUDAParticleSimulator::CCUDAParticleSimulator(CScene *scene, QVector3D boxSize, int device, SimulationScenario scenario, QObject *parent)
    : CBaseParticleSimulator(scene, boxSize, scenario, parent), m_device(device) {}

CCUDAParticleSimulator::~CCUDAParticleSimulator() = default;

QString CCUDAParticleSimulator::getSelectedDevice() { return CUDAPlatforms::getDeviceInfo(m_device); }

void CCUDAParticleSimulator::setGravityVector(QVector3D newGravity) {
    CBaseParticleSimulator::setGravityVector(newGravity);
    if (m_cuda) m_cuda->check(sph_set_gravity(m_cuda->ctx(), gravity.x(), gravity.y(), gravity.z()), "setGravityVector");
}

void CCUDAParticleSimulator::enableSlab(int rank, int world, const unsigned char ncclId[128]) {
    m_slab = true;
    m_rank = rank;
    m_world = world;
    std::memcpy(m_ncclId, ncclId, 128);
    if (sph_slab_plan(m_grid->zRes(), world, rank, &m_z0, &m_z1) != SPH_OK) throw CUDAException(sph_last_error(nullptr));
    setOwnedLayers(m_z0, m_z1);
}

void CCUDAParticleSimulator::setupScene() {
    CBaseParticleSimulator::setupScene();  // fills m_clParticles / m_maxParticlesCount

    // ≙ CGPUBaseParticleSimulator::setupKernels: size the device buffers, hand over walls and constants
    // device capacity: the whole scene on one device; in slab mode the owned share plus room for the ghost
    // layers and for load drifting between slabs
    uint32_t capacity = m_maxParticlesCount > 0 ? m_maxParticlesCount : 1;
    if (m_slab) {
        const uint64_t face = (uint64_t)m_grid->xRes() * m_grid->yRes() * 4 * 48;
        capacity = (uint32_t)std::min<uint64_t>((uint64_t)m_particlesCount * 13 / 10 + 2 * face + 1024, 0x7ffffff0u);
    }
    sph_config cfg;
    sph_config_init(&cfg, m_boxSize.x(), m_boxSize.y(), m_boxSize.z(), capacity);
    cfg.grid_res[0] = m_grid->xRes();
    cfg.grid_res[1] = m_grid->yRes();
    cfg.grid_res[2] = m_grid->zRes();
    cfg.dt = dt;
    cfg.gravity[0] = gravity.x();
    cfg.gravity[1] = gravity.y();
    cfg.gravity[2] = gravity.z();
    const auto &walls = m_grid->getCollisionGeometry()->getBoundingBox().m_walls;
    cfg.wall_count = (int32_t)walls.size();
    std::memcpy(cfg.walls, walls.data(), walls.size() * sizeof(sWall));
    cfg.device = m_device;
    if (m_slab) {
        cfg.rank = m_rank;
        cfg.world = m_world;
        std::memcpy(cfg.nccl_id, m_ncclId, 128);
    }
    m_cuda.reset(new CUDAWrapper(cfg, m_slab));

    // page-lock the host mirror once: it never reallocates (reserved to the maximum count in setupScene; in slab mode
    // to the rank's device capacity, because the owned count drifts as particles migrate)
    if (m_slab) m_clParticles.reserve(capacity);
    if (m_clParticles.capacity() > 0)
        sph_pin_host_buffer(m_cuda->ctx(), m_clParticles.data(), m_clParticles.capacity() * sizeof(CParticle::Physics));

    pushCollisionFaces();
    pushEmitter();

    m_deviceCount = 0;
    m_cuda->check(sph_upload_particles(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(m_clParticles.data()),
                                       (uint32_t)m_particlesCount), "setupScene upload");
    m_deviceCount = (cl_uint)m_particlesCount;
}

void CCUDAParticleSimulator::setCollisionFaces(const std::vector<sFace> &faces) {
    m_grid->getCollisionGeometry()->setFaces(faces);
    if (m_cuda) pushCollisionFaces();
}

void CCUDAParticleSimulator::pushCollisionFaces() {
    const auto &faces = m_grid->getCollisionGeometry()->getFaces();
    std::vector<sph_face> flat(faces.size());
    for (size_t f = 0; f < faces.size(); ++f) {
        const sVertex *verts[3] = {&faces[f].m_v0, &faces[f].m_v1, &faces[f].m_v2};
        float *dst[3] = {flat[f].v0, flat[f].v1, flat[f].v2};
        flat[f].normal[0] = faces[f].m_normal.x();
        flat[f].normal[1] = faces[f].m_normal.y();
        flat[f].normal[2] = faces[f].m_normal.z();
        for (int v = 0; v < 3; ++v) {
            dst[v][0] = verts[v]->m_pos.x();
            dst[v][1] = verts[v]->m_pos.y();
            dst[v][2] = verts[v]->m_pos.z();
        }
    }
    m_cuda->check(sph_set_collision_faces(m_cuda->ctx(), flat.data(), (uint32_t)flat.size()), "collision faces");
}

void CCUDAParticleSimulator::setEmissionMultiplier(int nozzles) {
    CBaseParticleSimulator::setEmissionMultiplier(nozzles);
    if (m_cuda) pushEmitter();
}

void CCUDAParticleSimulator::pushEmitter() {
    // generateParticles() as device templates (fountain only): stepMany() then emits without touching PCIe
    if (m_slab) return;
    const std::vector<CParticle::Physics> tpl = emissionTemplate();
    m_cuda->check(sph_set_emitter(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(tpl.data()), (uint32_t)tpl.size(), 7,
                                  m_maxParticlesCount), "emitter");
}

// Fountain batches: the device appends the new particles itself (sph_set_emitter) and runs the fused step; the host
// applies the same emission rule to its mirror, so ids and counts stay in lock-step without any per-step upload.
void CCUDAParticleSimulator::stepManyFountain(int steps, double *deviceMs) {
    pushNewParticles();  // anything step() emitted on the host but has not sent yet
    m_cuda->check(sph_step(m_cuda->ctx(), steps, deviceMs), "stepMany (fountain)");
    for (int k = 0; k < steps; ++k) emitParticles();
    m_deviceCount = (cl_uint)m_particlesCount;
    uint32_t n = 0;
    sph_particle_count(m_cuda->ctx(), &n);
    if (n != m_deviceCount) throw CUDAException("fountain: device and host emission counts disagree");
    addIterations((unsigned long)steps);
}

void CCUDAParticleSimulator::pushNewParticles() {
    // fountain: generateParticles() appended to the host mirror; send only the new tail
    if ((cl_uint)m_particlesCount > m_deviceCount) {
        m_cuda->check(sph_append_particles(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(m_clParticles.data()) + m_deviceCount,
                                           (uint32_t)m_particlesCount - m_deviceCount), "append");
        m_deviceCount = (cl_uint)m_particlesCount;
    }
}

void CCUDAParticleSimulator::step() {
    if (m_slab) {  // slab mode: the exchange is part of the fused device step
        try {
            // RoundTrip: the host vector (this rank's owned particles, ids in the records) is canonical and goes up
            // before every step, like on a single device
            if (m_mirrorMode == RoundTrip)
                m_cuda->check(sph_upload_particles(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(m_clParticles.data()),
                                                   (uint32_t)m_clParticles.size()), "step upload");
            m_cuda->check(sph_step(m_cuda->ctx(), 1, nullptr), "step");
            if (m_mirrorMode != Resident) syncHostMirror();
        } catch (CUDAException &exc) {
            emitErrorOccured(exc.what());
            stop();
        }
        return;
    }
    try {
        if (m_mirrorMode == RoundTrip && m_cuda) {
            m_cuda->check(sph_upload_particles(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(m_clParticles.data()),
                                               m_deviceCount), "step upload");
        }
        if (m_scenario == DAM_BREAK && !m_brute && !sampleThisStep()) {
            // nothing to emit and no phase durations wanted for this step: the five phases as ONE fused device step
            // (same bits as the phase path; the phase calls remain for sampled steps, the fountain and all-pairs mode)
            m_cuda->check(sph_step(m_cuda->ctx(), 1, nullptr), "step");
        } else {
            CBaseParticleSimulator::step();
        }
        const bool refresh = (getTotalIteration() + 1) % (unsigned long)m_mirrorStride == 0;
        if (m_mirrorMode == RoundTrip || (m_mirrorMode == Download && refresh)) {
            syncHostMirror();
        } else if (m_mirrorMode == AsyncDownload && refresh) {
            // snapshot now (device-side, ~15 us per million particles), PCIe copy overlapped with the next steps
            m_cuda->check(sph_download_particles_async(m_cuda->ctx(), reinterpret_cast<sph_particle *>(m_clParticles.data()),
                                                       (uint32_t)m_clParticles.size()), "async mirror");
        }
    } catch (CUDAException &exc) {
        emitErrorOccured(exc.what());
        stop();
    }
}

void CCUDAParticleSimulator::stepMany(int steps, double *deviceMs) {
    if (!m_cuda) throw CUDAException("stepMany before setupScene");
    // The mirror modes keep their meaning across a batch: in RoundTrip the host vector is canonical, so it goes up
    // first; in every non-resident mode the mirror shows the state after the batch (one read-back, not one per step).
    if (m_mirrorMode == RoundTrip)
        m_cuda->check(sph_upload_particles(m_cuda->ctx(), reinterpret_cast<const sph_particle *>(m_clParticles.data()),
                                           m_slab ? (uint32_t)m_clParticles.size() : m_deviceCount), "stepMany upload");
    if (m_brute && !m_slab) {  // all-pairs semantics go through the phase path
        for (int k = 0; k < steps; ++k) { CBaseParticleSimulator::step(); addIterations(1); }
        if (deviceMs) { m_cuda->check(sph_synchronize(m_cuda->ctx()), "stepMany"); *deviceMs = 0.0; }
    } else if (m_scenario == FOUNTAIN && !m_slab) {
        stepManyFountain(steps, deviceMs);
    } else {
        m_cuda->check(sph_step(m_cuda->ctx(), steps, deviceMs), "stepMany");
        addIterations((unsigned long)steps);
    }
    if (m_mirrorMode != Resident && steps > 0) syncHostMirror();
}

void CCUDAParticleSimulator::setMirrorMode(MirrorMode m) {
    // RoundTrip makes the host vector canonical (it is uploaded before every step, like the reference's OpenCL
    // path).  Entering it after resident steps with a stale mirror would silently rewind the simulation to whatever
    // the mirror last held, so the mirror is brought up to date first.
    if (m == RoundTrip && m_mirrorMode != RoundTrip && m_cuda) syncHostMirror();
    m_mirrorMode = m;
}

void CCUDAParticleSimulator::waitHostMirror() {
    if (m_cuda) m_cuda->check(sph_download_wait(m_cuda->ctx(), nullptr), "waitHostMirror");
}

void CCUDAParticleSimulator::syncHostMirror() {
    uint32_t n = 0;
    if (m_slab) {  // owned particles only, compact, canonical order
        sph_particle_count(m_cuda->ctx(), &n);
        m_clParticles.resize(n, CParticle::Physics(0, 0, 0, 0));
        m_particlesCount = (cl_int)n;
        m_cuda->check(sph_download_owned(m_cuda->ctx(), reinterpret_cast<sph_particle *>(m_clParticles.data()), n, &n), "syncHostMirror");
        return;
    }
    m_cuda->check(sph_download_particles(m_cuda->ctx(), reinterpret_cast<sph_particle *>(m_clParticles.data()),
                                         (uint32_t)m_clParticles.size(), &n), "syncHostMirror");
}

// Phase durations are only measured (which forces a device sync) on the steps the profiler samples;
// otherwise the phases are enqueued asynchronously and report 0, like the reference with PROFILING off.
double CCUDAParticleSimulator::updateGrid() {
    pushNewParticles();
    double ms = 0.0;
    if (m_brute) return 0.0;  // ≙ CGPUBruteParticleSimulator::updateGrid: nothing to build
    m_cuda->check(sph_update_grid(m_cuda->ctx(), sampleThisStep() ? &ms : nullptr), "updateGrid");
    return ms;
}

double CCUDAParticleSimulator::updateDensityPressure() {
    double ms = 0.0;
    double *out = sampleThisStep() ? &ms : nullptr;
    m_cuda->check(m_brute ? sph_brute_density_pressure(m_cuda->ctx(), out) : sph_density_pressure(m_cuda->ctx(), out),
                  "updateDensityPressure");
    return ms;
}

double CCUDAParticleSimulator::updateForces() {
    double ms = 0.0;
    double *out = sampleThisStep() ? &ms : nullptr;
    m_cuda->check(m_brute ? sph_brute_forces(m_cuda->ctx(), out) : sph_forces(m_cuda->ctx(), out), "updateForces");
    return ms;
}

double CCUDAParticleSimulator::updateCollisions() {
    double ms = 0.0;
    m_cuda->check(sph_collisions(m_cuda->ctx(), &ms), "updateCollisions");
    return ms;
}

double CCUDAParticleSimulator::integrate() {
    double ms = 0.0;
    m_cuda->check(sph_integrate(m_cuda->ctx(), sampleThisStep() ? &ms : nullptr), "integrate");
    return ms;
}
