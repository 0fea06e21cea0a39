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
    //   AsyncDownload — like Download, but the read-back runs on a copy stream while the next steps compute
    //               (sph_download_particles_async); waitHostMirror() completes the snapshot in flight, so the mirror
    //               shows the state of the last refresh step without ever stalling the simulation on PCIe
    enum MirrorMode { Resident = 0, Download = 1, RoundTrip = 2, AsyncDownload = 3 };

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
    void syncHostMirror();                                 // device -> m_clParticles (indexed by id), blocking, current state
    void waitHostMirror();                                 // AsyncDownload: finish the read-back in flight (no-op otherwise)
    void setMirrorMode(MirrorMode m);
    // viewer bridge: in Download mode refresh the host mirror only every `stride`-th step (a 60 Hz viewer does not
    // need 1500 read-backs per second); 1 = every step like the reference's OpenCL path
    void setMirrorStride(int stride) { m_mirrorStride = stride < 1 ? 1 : stride; }
    // collision mesh for CCollisionGeometry::inverseBounce (row f4); may be called before or after setupScene
    void setCollisionFaces(const std::vector<sFace> &faces);
    void setEmissionMultiplier(int nozzles) override;       // also re-arms the device-side emitter
    void setBruteForce(bool on) { m_brute = on; }          // CGPUBruteParticleSimulator semantics (all pairs)
    // Multi-GPU extension: make this instance rank `rank` of `world` z-slabs (call before setupScene).  ncclId is
    // the 128-byte id from sph_comm_unique_id(), identical on all ranks.  In slab mode the host mirror holds this
    // rank's owned particles only (compact, ids in the records) and step() runs the fused device step.
    void enableSlab(int rank, int world, const unsigned char ncclId[128]);
    bool slabMode() const { return m_world > 1 || m_slab; }
    void slabRange(int &z0, int &z1) const { z0 = m_z0; z1 = m_z1; }
    sph_context *context() const { return m_cuda ? m_cuda->ctx() : nullptr; }

protected:
    double updateGrid() override;
    double updateDensityPressure() override;
    double updateForces() override;
    double updateCollisions() override;
    double integrate() override;

private:
    void pushNewParticles();
    void pushCollisionFaces();
    void pushEmitter();
    void stepManyFountain(int steps, double *deviceMs);

    bool m_slab = false;
    int m_rank = 0, m_world = 1, m_z0 = 0, m_z1 = 0;
    unsigned char m_ncclId[128] = {0};

    int m_device;
    std::unique_ptr<CUDAWrapper> m_cuda;
    MirrorMode m_mirrorMode = Resident;
    int m_mirrorStride = 1;
    bool m_brute = false;
    cl_uint m_deviceCount = 0;  // particles already on the device
};
