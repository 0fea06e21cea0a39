// QtCompat.h — the handful of Qt types the simulator interface mentions, as plain C++.
//
// The reference's CBaseParticleSimulator is a QObject whose signature uses QVector3D, QString,
// QList, qint64, Qt::Key, QTimer and QElapsedTimer (include/CBaseParticleSimulator.h:4-11,29-110).
// The headless core keeps those names so simulator code reads the same, but nothing here links Qt:
// when the optional Qt3D viewer is built, define SPH_WITH_QT and the real headers are used instead.
#pragma once

#ifdef SPH_WITH_QT
#include <QElapsedTimer>
#include <QList>
#include <QObject>
#include <QString>
#include <QTimer>
#include <QVector3D>
#else

#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

using qint64 = long long;
using QString = std::string;

namespace Qt {
enum Key { Key_Space = 0x20, Key_G = 0x47, Key_O = 0x4f, Key_P = 0x50, Key_R = 0x52, Key_S = 0x53 };
}

// Three floats with QVector3D's operator semantics (every operation rounds to fp32).
class QVector3D {
public:
    constexpr QVector3D() : m_v{0.f, 0.f, 0.f} {}
    constexpr QVector3D(float x, float y, float z) : m_v{x, y, z} {}
    float x() const { return m_v[0]; }
    float y() const { return m_v[1]; }
    float z() const { return m_v[2]; }
    void setX(float v) { m_v[0] = v; }
    void setY(float v) { m_v[1] = v; }
    void setZ(float v) { m_v[2] = v; }
    float lengthSquared() const { return m_v[0] * m_v[0] + m_v[1] * m_v[1] + m_v[2] * m_v[2]; }
    float length() const {
        const double s = double(m_v[0]) * m_v[0] + double(m_v[1]) * m_v[1] + double(m_v[2]) * m_v[2];
        return float(std::sqrt(s));
    }
    QVector3D &operator+=(const QVector3D &o) { m_v[0] += o.m_v[0]; m_v[1] += o.m_v[1]; m_v[2] += o.m_v[2]; return *this; }
    QVector3D &operator*=(float f) { m_v[0] *= f; m_v[1] *= f; m_v[2] *= f; return *this; }
    friend QVector3D operator+(const QVector3D &a, const QVector3D &b) { return {a.x() + b.x(), a.y() + b.y(), a.z() + b.z()}; }
    friend QVector3D operator-(const QVector3D &a, const QVector3D &b) { return {a.x() - b.x(), a.y() - b.y(), a.z() - b.z()}; }
    friend QVector3D operator-(const QVector3D &a) { return {-a.x(), -a.y(), -a.z()}; }
    friend QVector3D operator*(const QVector3D &a, float f) { return {a.x() * f, a.y() * f, a.z() * f}; }
    friend QVector3D operator*(float f, const QVector3D &a) { return {a.x() * f, a.y() * f, a.z() * f}; }
    friend QVector3D operator/(const QVector3D &a, float d) { return {a.x() / d, a.y() / d, a.z() / d}; }
    static float dotProduct(const QVector3D &a, const QVector3D &b) { return a.x() * b.x() + a.y() * b.y() + a.z() * b.z(); }

private:
    float m_v[3];
};

// QList<T> is only used as an append-only log (events << durations).
template <typename T>
class QList : public std::vector<T> {
public:
    QList &operator<<(const T &v) { this->push_back(v); return *this; }
};

class QObject {
public:
    explicit QObject(QObject *parent = nullptr) : m_parent(parent) {}
    virtual ~QObject() = default;
    QObject *parent() const { return m_parent; }

private:
    QObject *m_parent;
};

class QElapsedTimer {
public:
    void start() { m_t0 = clock::now(); }
    qint64 restart() { const qint64 e = elapsed(); start(); return e; }
    qint64 elapsed() const { return std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - m_t0).count(); }

private:
    using clock = std::chrono::steady_clock;
    clock::time_point m_t0 = clock::now();
};

// A 0-ms QTimer drives doWork() from the GUI event loop in the reference; headless callers drive
// step() themselves, so the timer only tracks whether the simulation counts as running.
class QTimer {
public:
    void start() { m_active = true; }
    void stop() { m_active = false; }
    bool isActive() const { return m_active; }

private:
    bool m_active = false;
};

#endif  // SPH_WITH_QT
