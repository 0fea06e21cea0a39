/*
 * ref_driver.cpp — C entry points over the REFERENCE'S OWN CPU simulator, compiled unmodified.
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile (target `ref`) compiles, from where they lie under
 * /root/reference and without touching them,
 *     src/CCPUParticleSimulator.cpp  src/CBaseParticleSimulator.cpp  src/CCollisionGeometry.cpp
 *     src/CGrid.cpp  src/CParticle.cpp  src/renderableentity.cpp  src/CScene.cpp
 * against the stand-in headers in oracle/qt_shim/ and links them with this file into
 * oracle/_ref/libsph_ref.so.  This file contains no simulation arithmetic: it only
 *   (1) supplies what Qt's moc / the renderer would (signal bodies, the wireframe material's ctor),
 *   (2) subclasses CCPUParticleSimulator to reach its protected members,
 *   (3) copies state in and out through plain C arrays indexed by particle id.
 * tests/test_ref_pins_oracle.py requires oracle/sph_oracle.cpp to reproduce this library bit for bit;
 * bench.py --impl reference times it ("kind": "reference").
 */
#include "CCPUParticleSimulator.h"
#include "CWireframeMaterial.h"

#include <cstdint>

/* ---- what moc would generate for the two signals (include/CBaseParticleSimulator.h:60-62) ---- */
static unsigned long g_last_iteration = 0;
static std::string g_last_error;
void CBaseParticleSimulator::iterationChanged(unsigned long iteration) { g_last_iteration = iteration; }
void CBaseParticleSimulator::errorOccured(const char *error) { g_last_error = error ? error : ""; }

/* ---- renderer-only pieces CGrid.cpp constructs (src/CGrid.cpp:41-42); shaders are not part of the path ---- */
CWireframeMaterial::CWireframeMaterial(Qt3DCore::QNode *parent)
    : Qt3DRender::QMaterial(parent), effect(nullptr), gl3Technique(nullptr), gl3Pass(nullptr), glShader(nullptr) {}
CWireframeMaterial::~CWireframeMaterial() {}

namespace {

class RefSim : public CCPUParticleSimulator
{
public:
    RefSim(CScene *scene, float box, SimulationScenario sc) : CCPUParticleSimulator(scene, box, sc) {}

    CGrid *grid() { return m_grid; }
    int64_t count() const { return (int64_t) m_particlesCount; }
    int64_t maxCount() const { return (int64_t) m_maxParticlesCount; }
    QVector3D gravityVector() const { return gravity; }
    QVector3D box() const { return m_boxSize; }
    float timeStep() const { return dt; }
    void params(float *out3) const
    {
        out3[0] = m_systemParams.poly6_constant;
        out3[1] = m_systemParams.spiky_constant;
        out3[2] = m_systemParams.viscosity_constant;
    }
    bool running() { return isRunning(); }

    template<typename F>
    void forEachParticle(F &&f)
    {
        std::vector<CParticle *> *cells = m_grid->getData();
        const int n = m_grid->getCellCount();
        for (int c = 0; c < n; ++c)
            for (CParticle *p : cells[c]) f(c, p);
    }
};

struct Ref
{
    CScene *scene;
    RefSim *sim;
};

}  // namespace

extern "C" {

void *ref_create(float box, int scenario)
{
    Ref *r = new Ref();
    r->scene = new CScene();
    r->sim = new RefSim(r->scene, box, scenario == 1 ? FOUNTAIN : DAM_BREAK);
    return r;
}

void ref_destroy(void *h)
{
    Ref *r = (Ref *) h;
    if (!r) return;
    delete r->sim;    /* ~CBaseParticleSimulator deletes m_grid */
    delete r->scene;  /* ~CScene deletes the root entity, which owns every CParticle (Qt parent/child ownership) */
    delete r;
}

void ref_setup_scene(void *h) { ((Ref *) h)->sim->setupScene(); }

/* CBaseParticleSimulator::step(), n times: emit -> grid -> density/pressure -> forces -> collisions -> integrate */
void ref_step(void *h, int n)
{
    RefSim *s = ((Ref *) h)->sim;
    for (int i = 0; i < n; ++i) s->step();
}

/* the five phases, public overrides of CCPUParticleSimulator (each returns QElapsedTimer integer ms) */
double ref_update_grid(void *h) { return ((Ref *) h)->sim->updateGrid(); }
double ref_update_density_pressure(void *h) { return ((Ref *) h)->sim->updateDensityPressure(); }
double ref_update_forces(void *h) { return ((Ref *) h)->sim->updateForces(); }
double ref_update_collisions(void *h) { return ((Ref *) h)->sim->updateCollisions(); }
double ref_integrate(void *h) { return ((Ref *) h)->sim->integrate(); }

void ref_set_gravity(void *h, float x, float y, float z) { ((Ref *) h)->sim->setGravityVector(QVector3D(x, y, z)); }
void ref_get_gravity(void *h, float *out3)
{
    QVector3D g = ((Ref *) h)->sim->gravityVector();
    out3[0] = g.x(); out3[1] = g.y(); out3[2] = g.z();
}
/* CBaseParticleSimulator::onKeyPressed (S, space, G, O, P) */
void ref_key(void *h, int key) { ((Ref *) h)->sim->onKeyPressed((Qt::Key) key); }
int ref_is_running(void *h) { return ((Ref *) h)->sim->running() ? 1 : 0; }
unsigned long ref_last_iteration(void) { return g_last_iteration; }
void ref_device_name(void *h, char *buf, int len)
{
    std::string s = ((Ref *) h)->sim->getSelectedDevice().toStdString();
    if (len > 0) { std::strncpy(buf, s.c_str(), (size_t) len - 1); buf[len - 1] = 0; }
}

int64_t ref_count(void *h) { return ((Ref *) h)->sim->count(); }
int64_t ref_max_count(void *h) { return ((Ref *) h)->sim->maxCount(); }
void ref_grid_res(void *h, int *res3)
{
    CGrid *g = ((Ref *) h)->sim->grid();
    res3[0] = g->xRes(); res3[1] = g->yRes(); res3[2] = g->zRes();
}
void ref_params(void *h, float *out3) { ((Ref *) h)->sim->params(out3); }
float ref_dt(void *h) { return ((Ref *) h)->sim->timeStep(); }

/* the six walls CCollisionGeometry::init() extracted from the cuboid: normal xyz, position xyz per wall */
void ref_walls(void *h, float *out36)
{
    const sBoundingBox &bb = ((Ref *) h)->sim->grid()->getCollisionGeometry()->getBoundingBox();
    for (int w = 0; w < 6; ++w) {
        const sWall &wall = bb.m_walls[w];
        out36[6 * w + 0] = wall.normal.x; out36[6 * w + 1] = wall.normal.y; out36[6 * w + 2] = wall.normal.z;
        out36[6 * w + 3] = wall.position.x; out36[6 * w + 4] = wall.position.y; out36[6 * w + 5] = wall.position.z;
    }
}

/* CCollisionGeometry::inverseBoundingBoxBounce / inverseBounce on one (position, velocity) */
void ref_wall_bounce(void *h, const float *pos3, const float *vel3, float *out3)
{
    QVector3D p(pos3[0], pos3[1], pos3[2]), v(vel3[0], vel3[1], vel3[2]);
    QVector3D a = ((Ref *) h)->sim->grid()->getCollisionGeometry()->inverseBoundingBoxBounce(p, v);
    out3[0] = a.x(); out3[1] = a.y(); out3[2] = a.z();
}
void ref_mesh_bounce(void *h, const float *pos3, const float *vel3, float *out3)
{
    QVector3D p(pos3[0], pos3[1], pos3[2]), v(vel3[0], vel3[1], vel3[2]);
    QVector3D a = ((Ref *) h)->sim->grid()->getCollisionGeometry()->inverseBounce(p, v);
    out3[0] = a.x(); out3[1] = a.y(); out3[2] = a.z();
}

/* per-particle state, arrays indexed by particle id; what: 0 position, 1 velocity, 2 acceleration (3 floats each) */
void ref_get_vec(void *h, int what, float *out3n)
{
    ((Ref *) h)->sim->forEachParticle([&](int, CParticle *p) {
        const QVector3D &v = what == 0 ? p->position() : what == 1 ? p->velocity() : p->acceleration();
        float *o = out3n + 3 * (size_t) p->getId();
        o[0] = v.x(); o[1] = v.y(); o[2] = v.z();
    });
}
/* what: 0 density, 1 pressure */
void ref_get_scalar(void *h, int what, float *outn)
{
    ((Ref *) h)->sim->forEachParticle([&](int, CParticle *p) { outn[p->getId()] = what == 0 ? p->density() : p->pressure(); });
}

/* overwrite positions and velocities of the EXISTING particles (n must equal the current count): the state a
 * simulator would be in had it integrated to these values; cell lists are left as they are and the next
 * updateGrid() moves every particle to its cell, as after any integrate() */
int ref_set_state(void *h, int64_t n, const float *pos3n, const float *vel3n)
{
    RefSim *s = ((Ref *) h)->sim;
    if (n != s->count()) return -1;
    s->forEachParticle([&](int, CParticle *p) {
        const size_t i = 3 * (size_t) p->getId();
        p->position() = QVector3D(pos3n[i], pos3n[i + 1], pos3n[i + 2]);
        p->velocity() = QVector3D(vel3n[i], vel3n[i + 1], vel3n[i + 2]);
    });
    return 0;
}

/* the grid as updateGrid() left it, in the reference's own (history-dependent) order inside every cell:
 * cell_start[cells + 1], ids[n] */
void ref_get_cells(void *h, int32_t *cell_start, int32_t *ids)
{
    RefSim *s = ((Ref *) h)->sim;
    std::vector<CParticle *> *cells = s->grid()->getData();
    const int n = s->grid()->getCellCount();
    int32_t k = 0;
    for (int c = 0; c < n; ++c) {
        cell_start[c] = k;
        for (CParticle *p : cells[c]) ids[k++] = (int32_t) p->getId();
    }
    cell_start[n] = k;
}

}  /* extern "C" */
