// CCUDAParticleSimulator.h — the B200 simulator behind CBaseParticleSimulator.
//
// Drop-in for CCPUParticleSimulator / CGPUParticleSimulator (include/CCPUParticleSimulator.h:16,
// include/CGPUParticleSimulator.h:13): same constructor shape (scene, boxSize, device, scenario,
// parent), same overrides.  All particle state lives on the device between steps; the host mirror
// m_clParticles is refreshed on demand (syncHostMirror) or per step in one of the mirror modes.
#pragma once

#include <memory>

#include "CBaseParticleSimulator.h"
#include "CUDAWrapper.h"

class CCUDAParticleSimulator : public CBaseParticleSimulator {
public:
    // What happens to the host mirror m_clParticles around each step():
    //   Resident  — nothing (state stays in HBM; call syncHostMirror() when the viewer needs it)
    //   Download  — read back after every step (what a per-frame viewer needs)
    //   RoundTrip — upload before and read back after every step: the reference OpenCL path's
    //               semantics, where the host vector is canonical (src/CGPUParticleSimulator.cpp:61,
    //               src/CGPUBaseParticleSimulator.cpp:84)
    enum MirrorMode { Resident = 0, Download = 1, RoundTrip = 2 };

    explicit CCUDAParticleSimulator(CScene *scene, float boxSize, int device = 0, SimulationScenario scenario = DAM_BREAK,
                                    QObject *parent = nullptr);
    explicit CCUDAParticleSimulator(CScene *scene, QVector3D boxSize, int device = 0, SimulationScenario scenario = DAM_BREAK,
                                    QObject *parent = nullptr);
    ~CCUDAParticleSimulator() override;

    void setGravityVector(QVector3D newGravity) override;
    QString getSelectedDevice() override;
    void setupScene() override;
    void step() override;

    // extensions used by the headless bench and the tests
    void stepMany(int steps, double *deviceMs = nullptr);  // fused steps on the device (dam break: one CUDA graph)
    void syncHostMirror();                                 // device -> m_clParticles (indexed by id)
    void setMirrorMode(MirrorMode m) { m_mirrorMode = m; }
    void setBruteForce(bool on) { m_brute = on; }          // CGPUBruteParticleSimulator semantics (all pairs)
    sph_context *context() const { return m_cuda ? m_cuda->ctx() : nullptr; }

protected:
    double updateGrid() override;
    double updateDensityPressure() override;
    double updateForces() override;
    double updateCollisions() override;
    double integrate() override;

private:
    void pushNewParticles();

    int m_device;
    std::unique_ptr<CUDAWrapper> m_cuda;
    MirrorMode m_mirrorMode = Resident;
    bool m_brute = false;
    cl_uint m_deviceCount = 0;  // particles already on the device
};
