"""Row f4: CCollisionGeometry::inverseBounce (src/CCollisionGeometry.cpp:97-115), the per-face bounce the reference
prepared for "collisions with a general object" and never calls.  CPU part: known answers derived from the source
and bit-for-bit agreement of the two restatements; GPU part: the CUDA path against the oracle through the C ABI."""
import numpy as np
import pytest

from np_restatement import NpSim
from oracle_binding import Oracle


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def face(normal, v0, v1, v2):
    return np.array([*normal, *v0, *v1, *v2], dtype=np.float32)


def ramp(box, c=0.04):
    """A chamfer across the bottom -x edge of the box (under the dam-break column): the plane (x+b) + (y+b) = c as two
    triangles.  The normal points outwards like the walls' normals and is deliberately not unit length."""
    b = box / 2
    n = (-0.5, -0.5, 0.0)
    lo, hi = (-b + c, -b), (-b, -b + c)
    return np.stack([
        face(n, (*lo, -b), (*hi, -b), (*lo, b)),
        face(n, (*hi, b), (*lo, b), (*hi, -b)),
    ])


def lone_particle(box, p, v=(0, 0, 0)):
    o = Oracle(box)
    o.set_state(np.array([p], dtype=np.float32), np.array([v], dtype=np.float32))
    return o


def forces_of(o):
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    return o.acc.astype(np.float64)


def test_known_answer_axis_plane():
    # a face whose three vertices share x = 0.1 and whose normal is +x acts like a right wall at 0.1: a particle at
    # rest s = 0.004 in front of it gets 3 vertices x 5000 x (0.01 - s) in -x on top of the shipped acceleration
    p = (0.1 - 0.004, 0.0, 0.0)
    base = forces_of(lone_particle(0.9, p))[0]
    o = lone_particle(0.9, p)
    o.set_faces([face((1, 0, 0), (0.1, -0.2, 0), (0.1, 0.2, 0), (0.1, 0, 0.3))])
    got = forces_of(o)[0] - base
    assert got[0] == pytest.approx(-3 * 5000.0 * 0.006, rel=2e-5) and got[1] == 0 and got[2] == 0
    # out of reach (d <= 0): nothing is added
    far = lone_particle(0.9, (0.1 - 0.02, 0.0, 0.0))
    far.set_faces([face((1, 0, 0), (0.1, -0.2, 0), (0.1, 0.2, 0), (0.1, 0, 0.3))])
    assert np.array_equal(forces_of(far), forces_of(lone_particle(0.9, (0.1 - 0.02, 0.0, 0.0))))


def test_normal_is_normalised_and_damping_acts_along_it():
    # normal (0, -2, 0): inverse (0, 2, 0) is normalised to (0, 1, 0), so the spring is 5000 d, not 10000 d; a particle
    # moving into the plane at vy = -1 gets +0.9 per vertex from the damping term
    q = (0.0, -0.1, 0.0)
    f = [face((0, -2, 0), q, (0.1, -0.1, 0), (0, -0.1, 0.1))]
    p = (0.02, -0.1 + 0.003, 0.02)
    rest, moving = lone_particle(0.9, p), lone_particle(0.9, p, (0, -1, 0))
    rest.set_faces(f); moving.set_faces(f)
    a_rest = forces_of(rest)[0] - forces_of(lone_particle(0.9, p))[0]
    a_mov = forces_of(moving)[0] - forces_of(lone_particle(0.9, p, (0, -1, 0)))[0]
    assert a_rest[1] == pytest.approx(3 * 5000.0 * 0.007, rel=2e-5)
    assert a_mov[1] - a_rest[1] == pytest.approx(3 * 0.9, rel=1e-3)
    # a degenerate (zero) normal stays zero: Qt's normalize() returns early and the face has no effect
    z = lone_particle(0.9, p)
    z.set_faces([face((0, 0, 0), q, q, q)])
    assert np.array_equal(forces_of(z), forces_of(lone_particle(0.9, p)))


def test_empty_mesh_is_the_shipped_behaviour():
    a = Oracle(0.3).setup_scene()
    b = Oracle(0.3).setup_scene()
    b.set_faces(ramp(0.3)); b.set_faces(np.zeros((0, 12), np.float32))
    a.step(3); b.step(3)
    assert np.array_equal(bits(a.pos), bits(b.pos)) and np.array_equal(bits(a.vel), bits(b.vel))


@pytest.mark.parametrize("box,steps", [(0.2, 6), (0.3, 2)])
def test_restatements_agree_bitwise_with_a_mesh(box, steps):
    o = Oracle(box).setup_scene()
    s = NpSim(box).setup_dam_break()
    o.set_faces(ramp(box)); s.set_faces(ramp(box))
    for k in range(steps):
        o.update_grid(); s.update_grid()
        o.update_density_pressure(); s.density_pressure()
        o.update_forces(); s.forces()
        assert np.array_equal(bits(o.acc), bits(s.arrays()[2])), f"acceleration differs at step {k}"
        o.integrate(); s.integrate()
        pos, vel, _, _, _ = s.arrays()
        assert np.array_equal(bits(o.pos), bits(pos)) and np.array_equal(bits(o.vel), bits(vel))
    plain = Oracle(box).setup_scene()
    plain.step(steps)
    assert not np.array_equal(bits(plain.pos), bits(o.pos)), "the ramp must change the flow"


# ------------------------------------------------------------------------------------------------- GPU
RTOL = 1e-5


def make_ctx(gws, box, o, variant=1, mesh=None):
    ctx = gws.SphContext(box, o.n)
    ctx.set_option("neighbour_variant", variant)
    if mesh is not None:
        ctx.set_collision_faces(mesh)
    ctx.upload(gws.particles_from_arrays(o.pos, o.vel))
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 0])
def test_gpu_mesh_forces_match_oracle(gws, variant):
    box = 0.9
    o = Oracle(box).setup_scene()
    o.set_faces(ramp(box))
    o.step(12)  # the column is slumping onto the ramp
    pos, vel = o.pos, o.vel
    ctx = make_ctx(gws, box, o, variant, ramp(box))
    for sim in (o, ctx):
        sim.update_grid(); sim.density_pressure() if sim is ctx else sim.update_density_pressure()
    o.update_forces()
    ctx.forces()
    _, _, acc_sph = ctx.density_pressure_accel()
    ctx.collisions(); ctx.integrate()
    _, _, acc_tot = ctx.density_pressure_accel()
    assert (np.linalg.norm(o.acc_mesh, axis=1) > 0).sum() > 20, "the scene must exercise the mesh term"
    # wall and mesh terms come from identical pos/vel: total == fl(fl(sph_gpu + wall) + mesh) bit for bit
    assert np.array_equal(bits((acc_sph + o.acc_wall) + o.acc_mesh), bits(acc_tot))
    tol = RTOL * np.maximum(np.linalg.norm(o.acc, axis=1), o.acc_scale)
    assert np.all(np.abs(acc_tot - o.acc).max(axis=1) <= tol)
    # integrator: bit-exact given the device's own acceleration
    dt = np.float32(0.01)
    new_pos = (pos + vel * dt) + (acc_tot * dt) * dt
    rec = ctx.download()
    assert np.array_equal(bits(rec["position"][:, :3]), bits(new_pos))
    assert np.array_equal(bits(rec["velocity"][:, :3]), bits((new_pos - pos) / dt))
    ctx.close()


@pytest.mark.gpu
def test_gpu_mesh_fused_step_equals_phases_and_clears(gws):
    box = 0.9
    o = Oracle(box).setup_scene()
    a, b, c = make_ctx(gws, box, o, mesh=ramp(box)), make_ctx(gws, box, o, mesh=ramp(box)), make_ctx(gws, box, o)
    a.step(20)  # CUDA graph; forces + walls + mesh + integration in one kernel
    for _ in range(20):
        b.update_grid(); b.density_pressure(); b.forces(); b.collisions(); b.integrate()
    pa, pb = a.download(), b.download()
    assert np.array_equal(bits(pa["position"]), bits(pb["position"]))
    assert np.array_equal(bits(pa["velocity"]), bits(pb["velocity"]))
    c.step(20)
    pc = c.download()
    assert not np.array_equal(bits(pa["position"]), bits(pc["position"]))
    # changing the mesh mid-run re-captures the graph; clearing it restores the shipped behaviour bit for bit
    d = make_ctx(gws, box, o, mesh=ramp(box))
    d.step(1)
    d.set_collision_faces(np.zeros((0, 12), np.float32))
    d.upload(gws.particles_from_arrays(o.pos, o.vel))
    d.step(20)
    assert np.array_equal(bits(d.download()["position"]), bits(pc["position"]))
    # 20-step rollout against the oracle with the mesh: statistics (the flow is chaotic)
    o.set_faces(ramp(box))
    o.step(20)
    so, sg = o.stats(), a.stats()
    assert abs(sg["ke"] - so["ke"]) <= 1e-2 * so["ke"] and np.all(np.abs(sg["com"] - so["com"]) < 1e-3)
    for s in (a, b, c, d):
        s.close()


@pytest.mark.gpu
def test_gpu_mesh_through_the_simulator_facade(gws):
    box = 0.6
    sim = gws.Simulator("cuda", box)
    sim.set_collision_faces(ramp(box))  # before setupScene: kept by the collision geometry, pushed at setup
    sim.setup_scene()
    o = Oracle(box).setup_scene()
    o.set_faces(ramp(box))
    sim.set_mirror_mode(1)
    sim.step(5); o.step(5)
    hp = sim.host_particles()
    assert np.abs(hp["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    plain = Oracle(box).setup_scene()
    plain.step(5)
    assert np.abs(plain.pos - o.pos).max() > 1e-3 * 0.0457
    sim.close()
