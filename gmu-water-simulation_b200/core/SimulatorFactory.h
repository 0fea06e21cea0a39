// SimulatorFactory.h — the one-line touch point a caller needs: simulator type -> object.
// Mirrors MainWindow::createSimulator (src/mainwindow.cpp:171-201) and eSimulationType
// (include/mainwindow.h:60-63).  The reference's three values keep their numbers — the GUI uses them as
// combo-box row indices (src/mainwindow.cpp:90-97: insertItem((int) eSimulationType::GPUGrid, ...)) — and the
// CUDA types are APPENDED, so "GPU Grid" stays row 0 and "CUDA Grid" becomes row 3.  The CPU and OpenCL
// simulators are not part of this repository (no CPU fallback, no OpenCL), so asking for them throws.
#pragma once

#include <stdexcept>

#include "CCUDAParticleSimulator.h"

enum class eSimulationType { GPUGrid = 0, GPUBrute, CPU, CUDAGrid, CUDABrute };

// the combo-box texts (src/mainwindow.cpp:90-97) plus the two new rows; also the first field of the CSV logs
inline const char *simulationTypeName(eSimulationType type) {
    switch (type) {
        case eSimulationType::GPUGrid: return "GPU Grid";
        case eSimulationType::GPUBrute: return "GPU Brute Force";
        case eSimulationType::CPU: return "CPU Grid";
        case eSimulationType::CUDAGrid: return "CUDA Grid";
        case eSimulationType::CUDABrute: return "CUDA Brute Force";
    }
    return "";
}

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
