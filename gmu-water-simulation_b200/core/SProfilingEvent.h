// SProfilingEvent.h — per-step record of the five phase durations (milliseconds) plus fps.
// Same fields and meaning as the reference's sProfilingEvent (include/SProfilingEvent.h:8-31), so the
// CSV logs (src/mainwindow.cpp:310-368) written from these records stay comparable.
#pragma once

struct sProfilingEvent {
    unsigned long iteration = 0;
    double fps = 0.0;
    double updateGrid = 0.0;
    double updateDensityPressure = 0.0;
    double updateForces = 0.0;
    double updateCollisions = 0.0;
    double integrate = 0.0;

    explicit sProfilingEvent(unsigned long it = 0) : iteration(it) {}

    double together() const { return updateGrid + updateDensityPressure + updateForces + updateCollisions + integrate; }
};
