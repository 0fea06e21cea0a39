// SimulatorFactory.h — the one-line touch point a caller needs: simulator type -> object.
// Mirrors MainWindow::createSimulator (src/mainwindow.cpp:171-201) and eSimulationType
// (include/mainwindow.h:60-63) with the CUDA types added.  The CPU and OpenCL simulators are not part
// of this repository (no CPU fallback, no OpenCL), so asking for them throws.
#pragma once

#include <stdexcept>

#include "CCUDAParticleSimulator.h"

enum class eSimulationType { CPU = 0, GPUBrute, GPUGrid, CUDAGrid, CUDABrute };

inline CBaseParticleSimulator *createSimulator(eSimulationType type, CScene *scene, QVector3D boxSize, int device,
                                               SimulationScenario scenario) {
    switch (type) {
        case eSimulationType::CUDAGrid:
            return new CCUDAParticleSimulator(scene, boxSize, device, scenario);
        case eSimulationType::CUDABrute: {
            auto *sim = new CCUDAParticleSimulator(scene, boxSize, device, scenario);
            sim->setBruteForce(true);
            return sim;
        }
        default:
            throw std::invalid_argument("createSimulator: only the CUDA simulators are built in this repository");
    }
}
