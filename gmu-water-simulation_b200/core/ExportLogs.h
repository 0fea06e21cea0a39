// ExportLogs.h — the reference's profiling log format (row f3 of SURVEY.md §8).
// Same two CSV files and row layout as MainWindow::exportLogs (src/mainwindow.cpp:310-368):
//   <Scenario>_<box>.csv         one row per run : name;together(sample 0);together(sample 1);...
//   <Scenario>_<box>_detail.csv  five rows per run: name;<phase>;value;value;...
// so numbers logged by the CUDA simulator are directly comparable with the charts of the report.
#pragma once

#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>

#include "CBaseParticleSimulator.h"

inline void exportLogs(CBaseParticleSimulator &sim, const std::string &dir, const std::string &simName) {
    const std::string scenario = sim.m_scenario == DAM_BREAK ? "Dam_break" : "Fountain";
    char box[32];
    std::snprintf(box, sizeof(box), "%g", (double)sim.getBoxSize().x());
    const std::string base = dir + "/" + scenario + "_" + box;
    std::ofstream total(base + ".csv", std::ios::app), detail(base + "_detail.csv", std::ios::app);
    if (!total || !detail) throw std::runtime_error("exportLogs: cannot open " + base + ".csv");
    total << simName;
    for (const auto &e : sim.events) total << ";" << e.together();
    total << "\n";
    struct Row {
        const char *name;
        double sProfilingEvent::*field;
    };
    const Row rows[] = {{"Grid", &sProfilingEvent::updateGrid},
                        {"Density + pressure", &sProfilingEvent::updateDensityPressure},
                        {"Forces", &sProfilingEvent::updateForces},
                        {"Collisions", &sProfilingEvent::updateCollisions},
                        {"Integrate", &sProfilingEvent::integrate}};
    for (const Row &r : rows) {
        detail << simName << ";" << r.name;
        for (const auto &e : sim.events) detail << ";" << e.*(r.field);
        detail << "\n";
    }
}
