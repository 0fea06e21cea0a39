// CBaseParticleSimulator.cpp — scene setup, emission, phase sequencing (headless).
// Behaviour follows src/CBaseParticleSimulator.cpp of the reference; the fp32/fp64 mix of the lattice
// loops is kept exactly because step-0 neighbour sets depend on the accumulated fp32 coordinates.
#include "CBaseParticleSimulator.h"

#include <cassert>
#include <cmath>

CBaseParticleSimulator::CBaseParticleSimulator(CScene *scene, float boxSize, SimulationScenario scenario, QObject *parent)
    : CBaseParticleSimulator(scene, QVector3D(boxSize, boxSize, boxSize), scenario, parent) {}

CBaseParticleSimulator::CBaseParticleSimulator(CScene *scene, QVector3D boxSize, SimulationScenario scenario, QObject *parent)
    : QObject(parent), m_scenario(scenario), m_scene(scene), gravity(0, GRAVITY_ACCELERATION, 0), dt(0.01f),
      m_grid(nullptr), m_boxSize(boxSize), m_surfaceThreshold(0.01f) {
    // kernel constants in fp32, as the reference stores them for its device code (:23-25)
    const double h = CParticle::h;
    m_systemParams.poly6_constant = (cl_float)(315.0f / (64.0f * M_PI * std::pow(h, 9)));
    m_systemParams.spiky_constant = (cl_float)(-45.0f / (M_PI * std::pow(h, 6)));
    m_systemParams.viscosity_constant = (cl_float)(45.0f / (M_PI * std::pow(h, 6)));

    // ceil(box / h) per axis with a float division (:27-31)
    const QVector3D resolution((float)(int)std::ceil(m_boxSize.x() / CParticle::h),
                               (float)(int)std::ceil(m_boxSize.y() / CParticle::h),
                               (float)(int)std::ceil(m_boxSize.z() / CParticle::h));
    m_grid = new CGrid(m_boxSize, resolution);
}

void CBaseParticleSimulator::setupScene() {
    // :38-65.  halfParticle is a double holding the fp32 value h/2; the loop variables are fp32
    // accumulators ("y += halfParticle" rounds to fp32 on every add).
    const double halfParticle = CParticle::h / 2.0f;
    const unsigned int calculatedCount =
        (unsigned)(std::ceil(m_boxSize.z() / halfParticle) * std::ceil(m_boxSize.y() / halfParticle) *
                   std::ceil(m_boxSize.x() / 4 / halfParticle));
    if (!filtersScene()) m_clParticles.reserve(calculatedCount);

    if (m_scenario == DAM_BREAK) {
        const QVector3D offset = -m_boxSize / 2.0f;
        for (float y = 0; y < m_boxSize.y(); y += halfParticle)
            for (float x = 0; x < m_boxSize.x() / 4.0; x += halfParticle)
                for (float z = 0; z < m_boxSize.z(); z += halfParticle)
                    addParticle(x + offset.x(), y + offset.y(), z + offset.z());
        // The reference asserts calculatedCount == m_particlesCount (:56); that holds for its cubes.  A long
        // non-cubic tank (our extension) accumulates fp32 rounding over thousands of "z += halfParticle" and may
        // end up one lattice plane off the ceil() formula, so the identity is only asserted for cubes.
        assert(calculatedCount == m_nextParticleId || m_boxSize.x() != m_boxSize.z() || m_boxSize.x() != m_boxSize.y());
        (void)calculatedCount;
        m_maxParticlesCount = m_nextParticleId;
    } else {
        m_maxParticlesCount = calculatedCount;
    }
}

bool CBaseParticleSimulator::ownsParticle(float z) const {
    if (!filtersScene()) return true;
    // the z-layer exactly as updateGrid computes it (src/CCPUParticleSimulator.cpp:48,64-70)
    int layer = (int)std::floor(((double)z + (double)m_boxSize.z() / 2.0) / (double)CParticle::h);
    if (layer < 0) layer = 0;
    else if (layer >= m_grid->zRes()) layer = m_grid->zRes() - 1;
    return layer >= m_ownedZ0 && layer < m_ownedZ1;
}

void CBaseParticleSimulator::addParticle(float x, float y, float z, cl_float3 initialVelocity) {
    // :67-74 without the per-particle Qt3D entity; the id is the global running count
    const cl_uint id = m_nextParticleId++;
    if (!ownsParticle(z)) return;
    m_clParticles.emplace_back(x, y, z, id, initialVelocity);
    m_particlesCount++;
}

void CBaseParticleSimulator::nozzlePattern(int nozzle, float out[7][3]) const {
    // :195-206 — the seven positions one nozzle emits, on the box floor.  One nozzle sits at the origin like the
    // reference's; with more nozzles (extension) they form a centred square array 3h apart so their particles never
    // coincide.
    const float halfParticle = CParticle::h / 2.0f;
    const QVector3D offset = -m_boxSize / 2.0f;
    const int side = (int)std::ceil(std::sqrt((double)m_emissionMultiplier));
    const float cx = 3.0f * CParticle::h * ((float)(nozzle % side) - 0.5f * (float)(side - 1));
    const float cz = 3.0f * CParticle::h * ((float)(nozzle / side) - 0.5f * (float)(side - 1));
    const float dx[7] = {0, -halfParticle, halfParticle, -CParticle::h / 4, CParticle::h / 4, -CParticle::h / 4, CParticle::h / 4};
    const float dz[7] = {0, 0, 0, -halfParticle, -halfParticle, halfParticle, halfParticle};
    for (int k = 0; k < 7; ++k) {
        out[k][0] = cx + dx[k];
        out[k][1] = offset.y();
        out[k][2] = cz + dz[k];
    }
}

std::vector<CParticle::Physics> CBaseParticleSimulator::emissionTemplate() const {
    std::vector<CParticle::Physics> tpl;
    if (m_scenario != FOUNTAIN) return tpl;
    const cl_float3 initialVelocity = {0.0f, m_boxSize.y() * 3.2f, 0.0f, 0.0f};
    for (int nozzle = 0; nozzle < m_emissionMultiplier; ++nozzle) {
        float p[7][3];
        nozzlePattern(nozzle, p);
        for (int k = 0; k < 7; ++k) tpl.emplace_back(p[k][0], p[k][1], p[k][2], 0u, initialVelocity);
    }
    return tpl;
}

void CBaseParticleSimulator::generateParticles() {
    // :187-210 — seven particles per nozzle per step at the box floor, shot upwards
    if (m_scenario != FOUNTAIN) return;
    const int particlesPerIteration = 7;
    const cl_float3 initialVelocity = {0.0f, m_boxSize.y() * 3.2f, 0.0f, 0.0f};
    for (int nozzle = 0; nozzle < m_emissionMultiplier; ++nozzle) {
        if (m_nextParticleId >= (m_maxParticlesCount - (cl_uint)particlesPerIteration)) return;
        float p[7][3];
        nozzlePattern(nozzle, p);
        for (int k = 0; k < 7; ++k) addParticle(p[k][0], p[k][1], p[k][2], initialVelocity);
    }
}

void CBaseParticleSimulator::start() {
    iterationSincePaused = 0;
    m_timer.start();
    m_elapsed_timer.start();
}

void CBaseParticleSimulator::stop() { m_timer.stop(); }

void CBaseParticleSimulator::toggleSimulation() {
    if (m_timer.isActive()) {
        m_timer.stop();
    } else {
        m_elapsed_timer.restart();
        start();
    }
}

void CBaseParticleSimulator::toggleGravity() {
    if (gravity.length() > 0.0)
        setGravityVector(QVector3D(0, 0, 0));
    else
        setGravityVector(QVector3D(0, GRAVITY_ACCELERATION, 0));
}

void CBaseParticleSimulator::setGravityVector(QVector3D newGravity) { gravity = newGravity; }

void CBaseParticleSimulator::step() {
    // :116-144 — emit, then the five phases; durations are logged every eventLoggerStride-th step
    sProfilingEvent durations(totalIteration);
    generateParticles();
    durations.updateGrid = updateGrid();
    durations.updateDensityPressure = updateDensityPressure();
    durations.updateForces = updateForces();
    durations.updateCollisions = updateCollisions();
    durations.integrate = integrate();
    if (sampleThisStep()) {
        durations.fps = getFps();
        events << durations;
    }
}

void CBaseParticleSimulator::doWork() {
    this->step();
    ++totalIteration;
    ++iterationSincePaused;
    emitIterationChanged(totalIteration);
}

void CBaseParticleSimulator::onKeyPressed(Qt::Key key) {
    // :154-179
    switch (key) {
        case Qt::Key_S: doWork(); break;
        case Qt::Key_Space: toggleSimulation(); break;
        case Qt::Key_G: toggleGravity(); break;
        case Qt::Key_O: gravity.setX(gravity.x() - 1); setGravityVector(gravity); break;
        case Qt::Key_P: gravity.setX(gravity.x() + 1); setGravityVector(gravity); break;
        default: break;
    }
}

double CBaseParticleSimulator::getFps() {
    const double elapsed = getElapsedTime() / 1000.0;
    return elapsed > 0 ? iterationSincePaused / elapsed : 0.0;
}
