// sph_bench.cpp — headless driver: builds a scene exactly like the GUI's "Setup" button would
// (src/mainwindow.cpp:271-287), then runs step() in a loop and reports particle-steps/s and the
// reference's own per-phase split (sProfilingEvent).  No Qt, no viewer.
//
//   sph_bench [--scenario dam_break|fountain] [--box B | --box3 X Y Z] [--steps K] [--warmup W]
//             [--device D] [--brute] [--phases] [--mirror 0|1|2|3] [--mirror-stride K] [--nozzles M] [--csv DIR]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/sph_host.h"
#include "ExportLogs.h"
#include "SimulatorFactory.h"

int main(int argc, char **argv) {
    std::string scenario = "dam_break", csv;
    float box[3] = {0.9f, 0.9f, 0.9f};
    int steps = 100, warmup = 10, device = 0, mirror = 0, mirror_stride = 1, nozzles = 1;
    bool brute = false, phases = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(2); }
            return argv[++i];
        };
        if (a == "--scenario") scenario = next("--scenario");
        else if (a == "--box") box[0] = box[1] = box[2] = (float)std::atof(next("--box"));
        else if (a == "--box3") { for (int k = 0; k < 3; ++k) box[k] = (float)std::atof(next("--box3")); }
        else if (a == "--steps") steps = std::atoi(next("--steps"));
        else if (a == "--warmup") warmup = std::atoi(next("--warmup"));
        else if (a == "--device") device = std::atoi(next("--device"));
        else if (a == "--mirror") mirror = std::atoi(next("--mirror"));  // 0 resident, 1 download, 2 round trip, 3 async download
        else if (a == "--mirror-stride") mirror_stride = std::atoi(next("--mirror-stride"));
        else if (a == "--nozzles") nozzles = std::atoi(next("--nozzles"));
        else if (a == "--csv") csv = next("--csv");
        else if (a == "--brute") brute = true;
        else if (a == "--phases") phases = true;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    const SimulationScenario sc = scenario == "fountain" ? FOUNTAIN : DAM_BREAK;
    try {
        CScene scene;
        std::unique_ptr<CBaseParticleSimulator> base(createSimulator(brute ? eSimulationType::CUDABrute : eSimulationType::CUDAGrid,
                                                                     &scene, QVector3D(box[0], box[1], box[2]), device, sc));
        auto *sim = static_cast<CCUDAParticleSimulator *>(base.get());
        sim->setEmissionMultiplier(nozzles);
        sim->setMirrorMode((CCUDAParticleSimulator::MirrorMode)mirror);
        sim->setMirrorStride(mirror_stride);
        sim->onErrorOccured([](const char *what) { std::fprintf(stderr, "error: %s\n", what); std::exit(1); });
        sim->setupScene();
        std::fprintf(stderr, "device: %s\nparticles: %lu (max %lu)\n", sim->getSelectedDevice().c_str(), sim->getParticlesCount(),
                     sim->getMaxParticlesCount());

        // the fountain emits on the device (sph_set_emitter) and runs the fused step like the dam break; only --phases,
        // the all-pairs variant and the mirror modes go through the five separate phase calls
        const bool phase_path = phases || brute || mirror != 0;
        auto run = [&](int k) {
            if (phase_path) for (int s = 0; s < k; ++s) sim->doWork();
            else sim->stepMany(k);
        };
        run(warmup);
        sph_synchronize(sim->context());
        sim->setProfiling(phases);
        sim->eventLoggerStride = 1;
        sim->start();
        const auto t0 = std::chrono::steady_clock::now();
        double particle_steps = 0;
        if (phase_path) {
            for (int s = 0; s < steps; ++s) { sim->doWork(); particle_steps += (double)sim->getParticlesCount(); }
        } else if (sc == FOUNTAIN) {  // the particle count grows while the scene fills
            for (int s = 0; s < steps; ++s) { sim->stepMany(1); particle_steps += (double)sim->getParticlesCount(); }
        } else {
            sim->stepMany(steps);
            particle_steps = (double)steps * (double)sim->getParticlesCount();
        }
        sim->waitHostMirror();  // mode 3: the last overlapped read-back belongs to the timed region
        sph_synchronize(sim->context());
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double ph[5] = {0, 0, 0, 0, 0};
        for (const auto &e : sim->events) {
            ph[0] += e.updateGrid; ph[1] += e.updateDensityPressure; ph[2] += e.updateForces;
            ph[3] += e.updateCollisions; ph[4] += e.integrate;
        }
        const double ne = sim->events.empty() ? 1.0 : (double)sim->events.size();
        std::printf("{\"scenario\": \"%s\", \"box\": [%g, %g, %g], \"particles\": %lu, \"steps\": %d, \"seconds\": %.6f, "
                    "\"particle_steps_per_s\": %.6e, \"ms_per_step\": %.4f, \"phase_ms\": {\"grid\": %.4f, \"density\": %.4f, "
                    "\"forces\": %.4f, \"collisions\": %.4f, \"integrate\": %.4f}, \"mirror\": %d, \"brute\": %s}\n",
                    scenario.c_str(), box[0], box[1], box[2], sim->getParticlesCount(), steps, sec, particle_steps / sec,
                    1e3 * sec / steps, ph[0] / ne, ph[1] / ne, ph[2] / ne, ph[3] / ne, ph[4] / ne, mirror, brute ? "true" : "false");
        if (!csv.empty()) exportLogs(*sim, csv, simulationTypeName(brute ? eSimulationType::CUDABrute : eSimulationType::CUDAGrid));  // needs --phases to have samples
    } catch (const std::exception &e) {
        std::fprintf(stderr, "fatal: %s\n", e.what());
        return 1;
    }
    return 0;
}
