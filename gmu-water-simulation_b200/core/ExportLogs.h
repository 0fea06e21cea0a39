// ExportLogs.h — the reference's profiling log format (row f3 of SURVEY.md §8), byte for byte.
//
// MainWindow::exportLogs (src/mainwindow.cpp:310-368) appends to two files named after the scenario combo-box
// text and the box size, "<Scenario>_<box>" with Scenario = "Dam break" | "Fountain" (src/mainwindow.cpp:103-107)
// and box = QString::number(slider / 10.0) (':313-316; 'g' format, 6 significant digits):
//
//   <Scenario>_<box>.csv          per run ONE line:   <sim type>;<together 0>;<together 1>;...;<together k>;\n
//   <Scenario>_<box>_detail.csv   per run:            <sim type>\n
//                                                      Grid;<v0>;<v1>;...;\n
//                                                      Density + pressure;<v0>;...;\n
//                                                      Forces;...;\n  Collisions;...;\n  Integrate;...;\n
//                                                      \n\n
//
// i.e. every value is FOLLOWED by ';' (':332-339), the detail block starts with the name on its own line (':350)
// and ends with two blank lines (':362).  Numbers are QString::number(double) = "%g".  Files are opened in
// append mode (':323), so logs written here and by the reference's GUI can share a file.
#pragma once

#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>

#include "CBaseParticleSimulator.h"

// QString::number(double): format 'g', precision 6
inline std::string qtNumber(double v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%g", v);
    return buf;
}

// "<Scenario>_<box>" — src/mainwindow.cpp:313-316
inline std::string exportLogsBaseName(const CBaseParticleSimulator &sim) {
    const std::string scenario = sim.m_scenario == DAM_BREAK ? "Dam break" : "Fountain";
    return scenario + "_" + qtNumber((double)sim.getBoxSize().x());
}

inline void exportLogs(CBaseParticleSimulator &sim, const std::string &dir, const std::string &simName) {
    const std::string base = dir + "/" + exportLogsBaseName(sim);
    std::ofstream data(base + ".csv", std::ios::app | std::ios::binary), details(base + "_detail.csv", std::ios::app | std::ios::binary);
    if (!data || !details) throw std::runtime_error("exportLogs: cannot open " + base + ".csv");
    std::string duration, detailText[5];
    for (const auto &event : sim.events) {
        duration += qtNumber(event.together()) + ';';
        detailText[0] += qtNumber(event.updateGrid) + ';';
        detailText[1] += qtNumber(event.updateDensityPressure) + ';';
        detailText[2] += qtNumber(event.updateForces) + ';';
        detailText[3] += qtNumber(event.updateCollisions) + ';';
        detailText[4] += qtNumber(event.integrate) + ';';
    }
    // ---- summary file
    data << simName << ';' << duration << '\n';
    // ---- detail file
    details << simName << '\n';
    const char *labels[] = {"Grid", "Density + pressure", "Forces", "Collisions", "Integrate"};
    for (int index = 0; index < 5; ++index) details << labels[index] << ';' << detailText[index] << '\n';
    details << "\n\n";
}
