// CBaseParticleSimulator.h — the simulator interface (the drop-in boundary), headless.
//
// Public and protected surface follows include/CBaseParticleSimulator.h:29-110 of the reference:
// same method names, argument meaning and phase contract (five protected pure virtuals, each
// returning elapsed milliseconds).  Differences, all additive:
//   * Qt types come from QtCompat.h; signals are callback lists (onIterationChanged/onErrorOccured);
//   * a second constructor takes a non-cubic box (needed for the 64M-particle tank);
//   * profiling is a runtime switch (the reference's compile-time PROFILING, include/config.h:4);
//   * the host mirror m_clParticles is filled by setupScene()/generateParticles() exactly as the
//     reference does, but no Qt3D entity is created per particle (viewer is optional).
#pragma once

#include <functional>
#include <vector>

#include "CGrid.h"
#include "CParticle.h"
#include "CScene.h"
#include "QtCompat.h"
#include "SProfilingEvent.h"

#define GRAVITY_ACCELERATION (-9.80665f)

enum SimulationScenario {
    DAM_BREAK = 0,
    FOUNTAIN,
};

namespace Qt3DExtras { class QSphereMesh; class QPhongMaterial; }

class CBaseParticleSimulator : public QObject {
public:
    explicit CBaseParticleSimulator(CScene *scene, float boxSize, SimulationScenario scenario = DAM_BREAK, QObject *parent = nullptr);
    // extension: non-cubic box (x, y, z extents)
    explicit CBaseParticleSimulator(CScene *scene, QVector3D boxSize, SimulationScenario scenario = DAM_BREAK, QObject *parent = nullptr);
    ~CBaseParticleSimulator() override { delete m_grid; }

    virtual void setupScene();
    virtual void setGravityVector(QVector3D newGravity);
    virtual QString getSelectedDevice() = 0;

    void start();
    void stop();
    virtual void step();
    void toggleSimulation();
    void toggleGravity();

    qint64 getElapsedTime() { return m_elapsed_timer.elapsed(); }
    double getFps();
    unsigned long getParticlesCount() { return (unsigned long)m_particlesCount; }
    unsigned long getMaxParticlesCount() { return m_maxParticlesCount; }
    QList<sProfilingEvent> events;
    int eventLoggerStride = 10;
    SimulationScenario m_scenario;

    // viewer-only members of the reference (include/CBaseParticleSimulator.h:58-59); null when headless
    Qt3DExtras::QSphereMesh *particle_mesh = nullptr;
    Qt3DExtras::QPhongMaterial *particle_material = nullptr;

    // signals -> callbacks
    void onIterationChanged(std::function<void(unsigned long)> cb) { m_iterationChanged.push_back(std::move(cb)); }
    void onErrorOccured(std::function<void(const char *)> cb) { m_errorOccured.push_back(std::move(cb)); }

    // slot
    virtual void onKeyPressed(Qt::Key key);

    // one timer tick of the reference (private slot doWork, src/CBaseParticleSimulator.cpp:146-152):
    // step() + iteration counters + iterationChanged.  Public so a headless loop can drive it.
    void doWork();

    // extensions
    void setProfiling(bool on) { m_profiling = on; }
    bool profiling() const { return m_profiling; }
    unsigned long getTotalIteration() const { return totalIteration; }
    QVector3D getBoxSize() const { return m_boxSize; }
    QVector3D getGravityVector() const { return gravity; }
    bool running() { return isRunning(); }
    const std::vector<CParticle::Physics> &getHostParticles() const { return m_clParticles; }
    // fountain: number of 7-particle nozzles fired per step (1 = the reference's behaviour)
    virtual void setEmissionMultiplier(int nozzles) { m_emissionMultiplier = nozzles < 1 ? 1 : nozzles; }
    int emissionMultiplier() const { return m_emissionMultiplier; }
    // the records ONE step of the fountain appends (ids unset), in emission order: 7 per nozzle
    std::vector<CParticle::Physics> emissionTemplate() const;
    // slab decomposition: keep only the particles of z-layers [z0, z1) when the scene is generated
    void setOwnedLayers(int z0, int z1) { m_ownedZ0 = z0; m_ownedZ1 = z1; }

protected:
    struct alignas(16) SystemParams {
        cl_float poly6_constant;
        cl_float spiky_constant;
        cl_float viscosity_constant;
    };

    CScene *m_scene;
    QVector3D gravity;

    cl_float dt;
    CGrid *m_grid;
    QVector3D m_boxSize;
    cl_float m_surfaceThreshold;
    cl_uint m_maxParticlesCount = 0;
    cl_int m_particlesCount = 0;
    SystemParams m_systemParams;
    std::vector<CParticle::Physics> m_clParticles;

    virtual double updateGrid() = 0;
    virtual double updateDensityPressure() = 0;
    virtual double updateForces() = 0;
    virtual double updateCollisions() = 0;
    virtual double integrate() = 0;

    // Extension for slab decomposition: scene generation walks the WHOLE scene (ids are the global running
    // count, as in the reference) but only keeps the particles whose z-layer lies in [m_ownedZ0, m_ownedZ1).
    bool ownsParticle(float z) const;
    bool filtersScene() const { return m_ownedZ1 >= 0; }
    int m_ownedZ0 = 0, m_ownedZ1 = -1;  // -1: no filter (own everything)
    cl_uint m_nextParticleId = 0;       // global running count (== m_particlesCount unless the scene is filtered)

    bool isRunning() { return m_timer.isActive(); }
    void emitIterationChanged(unsigned long it) { for (auto &cb : m_iterationChanged) cb(it); }
    void emitErrorOccured(const char *what) { for (auto &cb : m_errorOccured) cb(what); }
    // true when this step's phase durations will be logged (every eventLoggerStride-th iteration)
    bool sampleThisStep() const { return m_profiling && eventLoggerStride > 0 && totalIteration % (unsigned long)eventLoggerStride == 0; }
    void addIterations(unsigned long k) { totalIteration += k; iterationSincePaused += k; }
    // the emission at the top of step() on its own (host mirror only): for subclasses that emit on the device
    void emitParticles() { generateParticles(); }

private:
    void init(SimulationScenario scenario);
    void addParticle(float x, float y, float z, cl_float3 initialVelocity = {0, 0, 0, 0});
    void generateParticles();
    void nozzlePattern(int nozzle, float out[7][3]) const;

    QTimer m_timer;
    QElapsedTimer m_elapsed_timer;
    unsigned long iterationSincePaused = 0;
    unsigned long totalIteration = 0;
    bool m_profiling = false;
    int m_emissionMultiplier = 1;
    std::vector<std::function<void(unsigned long)>> m_iterationChanged;
    std::vector<std::function<void(const char *)>> m_errorOccured;
};
