"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): cell keys, canonical permutation, cell ranges and neighbour sets are
bit-exact; density, pressure, acceleration within rel 1e-5 in fp32 (pressure with the absolute floor
k*rho because p = k (rho - rho0) cancels; acceleration relative to max(|a|, sum of |terms|)); the
integrator is bit-exact given the device's own acceleration.
"""
import os

import numpy as np
import pytest

from oracle_binding import DAM_BREAK, FOUNTAIN, Oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def state_after(box, steps, scenario=DAM_BREAK):
    o = Oracle(box, scenario).setup_scene()
    if steps:
        o.step(steps)
    return o


def make_ctx(gws, box, pos, vel, cap=None, variant=1):
    ctx = gws.SphContext(box, cap or max(len(pos), 1))
    ctx.set_option("neighbour_variant", variant)  # 1 = bitmask passes (production), 0 = plain float4 walk
    if os.environ.get("SPH_TUNING"):
        ctx.set_option("tuning", int(os.environ["SPH_TUNING"]))  # experiment switches of the kernels under test
    ctx.upload(gws.particles_from_arrays(pos, vel))
    return ctx


def pos_tolerance(o):
    """One step from identical inputs: dx = a dt^2 carries the acceleration tolerance, plus a few ulps."""
    dt2 = 1e-4
    a_tol = RTOL * np.maximum(np.linalg.norm(o.acc, axis=1), o.acc_scale)
    return (a_tol * dt2 + 4 * np.spacing(np.abs(o.pos).max(axis=1).astype(np.float32)))[:, None]


def check_density(o, rho, prs):
    assert np.all(np.abs(rho - o.density) <= RTOL * np.abs(o.density)), np.abs(rho / o.density - 1).max()
    floor = np.maximum(np.abs(o.pressure), 3.0 * o.density)
    assert np.all(np.abs(prs - o.pressure) <= RTOL * floor)


def check_acc(ref, scale, got):
    tol = RTOL * np.maximum(np.linalg.norm(ref, axis=1), scale)
    err = np.abs(got - ref).max(axis=1)
    assert np.all(err <= tol), f"worst acc error/tol = {(err / tol).max():.3f}"


def assert_neighbour_sets(ctx, oc, ol, variant=1):
    """Neighbour sets against the oracle's (src/CCPUParticleSimulator.cpp:122-127), exact.  The sets that matter are
    the ones the production force pass consumes: the hit words stored by k_density_mask (+ the overflow list), decoded
    by sph_download_mask_neighbours.  The walk of the separate validation kernel is kept as a cross-check."""
    if variant == 1 and ctx.uses_mask_passes:
        mc, ml = ctx.neighbours(source="mask")
        assert np.array_equal(mc, oc), "production hit mask: neighbour counts differ from the oracle"
        assert np.array_equal(ml, ol), "production hit mask: neighbour sets differ from the oracle"
    gc, gl = ctx.neighbours()
    assert np.array_equal(gc, oc)
    assert np.array_equal(gl, ol)


def phase_parity(gws, o, box, variant=1):
    """One step of both implementations from the oracle's current state, checked phase by phase."""
    pos, vel = o.pos, o.vel
    ctx = make_ctx(gws, box, pos, vel, variant=variant)
    # ---- grid
    o.update_grid()
    ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    cs, ids = o.cells()
    assert np.array_equal(ctx.cell_start(), cs)
    assert np.array_equal(ctx.permutation().astype(np.int32), ids)
    # ---- density / pressure / neighbour sets
    o.update_density_pressure()
    ctx.density_pressure()
    rho, prs, _ = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    oc, ol = o.neighbours()
    assert_neighbour_sets(ctx, oc, ol, variant)
    # ---- forces (SPH part), then walls + integration
    o.update_forces()
    ctx.forces()
    _, _, acc_sph = ctx.density_pressure_accel()
    check_acc(o.acc_sph, o.acc_scale, acc_sph)
    ctx.collisions()
    ctx.integrate()
    _, _, acc_tot = ctx.density_pressure_accel()
    # wall term is computed from identical pos/vel: total == fl32(sph_gpu + wall_oracle) bit for bit
    assert np.array_equal((acc_sph + o.acc_wall).view(np.uint32), acc_tot.view(np.uint32))
    check_acc(o.acc, o.acc_scale, acc_tot)
    # integrator: bit-exact given the device's own acceleration (src/CCPUParticleSimulator.cpp:220-221)
    dt = np.float32(0.01)
    new_pos = (pos + vel * dt) + (acc_tot * dt) * dt
    new_vel = (new_pos - pos) / dt
    rec = ctx.download()
    assert np.array_equal(rec["id"], np.arange(o.n, dtype=np.uint32))
    assert np.array_equal(rec["position"][:, :3].view(np.uint32), new_pos.view(np.uint32))
    assert np.array_equal(rec["velocity"][:, :3].view(np.uint32), new_vel.view(np.uint32))
    # and against the oracle's own integration within the acceleration tolerance
    tol = pos_tolerance(o)
    o.integrate()
    assert np.all(np.abs(rec["position"][:, :3] - o.pos) <= tol)
    # the record also carries this step's density/pressure/cell id, like the reference's read-back
    assert np.array_equal(rec["cell_id"].astype(np.int32), ctx.keys())
    assert np.array_equal(rec["density"], rho)
    ctx.close()


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("box,steps", [(0.4, 0), (0.4, 1), (0.4, 10), (0.4, 100), (0.9, 0), (0.9, 25), (0.2, 30), (0.1, 5)])
def test_phase_parity_dam_break(gws, box, steps, variant):
    # step 0 is the knife-edge case: thousands of lattice pairs sit at r == h exactly;
    # boxes 0.1/0.2 have grids narrower than 4 cells (the library falls back to the plain walk there)
    phase_parity(gws, state_after(box, steps), box, variant)


@pytest.mark.parametrize("variant", [1, 0])
def test_phase_parity_non_cubic_box(gws, variant):
    box = (0.5, 0.3, 0.7)
    phase_parity(gws, state_after(box, 12), box, variant)


@pytest.mark.parametrize("steps", [1, 40, 150])
def test_phase_parity_fountain_state(gws, steps):
    """Per-phase parity on fountain states (particles emitted on the floor plane, strong wall forces)."""
    phase_parity(gws, state_after(0.4, steps, FOUNTAIN), 0.4)


def test_phase_parity_golden_fixture(gws):
    """Inputs and expected outputs from the committed fixture (no oracle run involved)."""
    g = np.load(os.path.join(GOLDEN, "oracle_phase_0p4_step10.npz"))
    ctx = make_ctx(gws, 0.4, g["pos"], g["vel"])
    ctx.update_grid()
    ctx.density_pressure()
    ctx.forces()
    assert np.array_equal(ctx.keys(), g["keys"])
    assert np.array_equal(ctx.cell_start(), g["cell_start"])
    assert np.array_equal(ctx.permutation().astype(np.int32), g["perm"])
    counts, _ = ctx.neighbours(lists=False)
    assert np.array_equal(counts, g["nb_counts"])
    rho, prs, acc = ctx.density_pressure_accel()
    assert np.all(np.abs(rho - g["density"]) <= RTOL * g["density"])
    check_acc(g["acc_sph"], g["acc_scale"], acc)


@pytest.mark.parametrize("name", ["ref_dam_break_0p4", "ref_dam_break_0p9", "ref_fountain_0p4"])
def test_parity_against_reference_golden(gws, name):
    """Vectors written by the REFERENCE'S OWN code (oracle/_ref, tests/golden/make_ref_golden.py): from the stored
    state before step k, one device step must give the reference's grid exactly and its density / pressure /
    acceleration / new positions within the north-star tolerance.  No oracle run on the value side; the oracle only
    supplies the conditioning scale of the acceleration tolerance."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    box = float(g["box"])
    for k in [int(v) for v in g["steps"]]:
        pos, vel = g[f"s{k}_pos_before"], g[f"s{k}_vel_before"]
        n_after = len(g[f"s{k}_pos"])
        ctx = gws.SphContext(box, max(n_after, 1))
        ctx.upload(gws.particles_from_arrays(pos, vel))
        if n_after > len(pos):                           # fountain: this step's emission (closed form, :191-206)
            hp, h4, y0 = np.float32(0.0457) / np.float32(2), np.float32(0.0457) / np.float32(4), -np.float32(box) / np.float32(2)
            e = np.array([[0, y0, 0], [-hp, y0, 0], [hp, y0, 0], [-h4, y0, -hp], [h4, y0, -hp], [-h4, y0, hp], [h4, y0, hp]], dtype=np.float32)
            ev = np.tile(np.array([[0, np.float32(box) * np.float32(3.2), 0]], dtype=np.float32), (7, 1))
            ctx.append(gws.particles_from_arrays(e, ev, ids=np.arange(len(pos), len(pos) + 7)))
            pos, vel = np.vstack([pos, e]), np.vstack([vel, ev])
        assert ctx.n == n_after
        ctx.update_grid(); ctx.density_pressure(); ctx.forces()
        cs_ref, ids_ref = g[f"s{k}_cell_start"], g[f"s{k}_ids"]
        assert np.array_equal(ctx.cell_start(), cs_ref)
        perm = ctx.permutation().astype(np.int32)
        # the reference's intra-cell order is its swap-and-pop history; membership per cell must be identical
        cell_of_slot = np.repeat(np.arange(len(cs_ref) - 1), np.diff(cs_ref))
        ref_sorted = ids_ref[np.lexsort((ids_ref, cell_of_slot))]
        assert np.array_equal(perm, ref_sorted)
        rho, prs, acc_sph = ctx.density_pressure_accel()
        rho_ref, prs_ref = g[f"s{k}_density"], g[f"s{k}_pressure"]
        assert np.all(np.abs(rho - rho_ref) <= RTOL * np.abs(rho_ref))
        assert np.all(np.abs(prs - prs_ref) <= RTOL * np.maximum(np.abs(prs_ref), 3.0 * rho_ref))
        ctx.integrate()
        _, _, acc = ctx.density_pressure_accel()
        o = Oracle(box).set_state(pos, vel)
        o.update_grid(); o.update_density_pressure(); o.update_forces()
        check_acc(g[f"s{k}_acc"], o.acc_scale, acc)
        rec = ctx.download()
        assert np.all(np.abs(rec["position"][:, :3] - g[f"s{k}_pos"]) <= pos_tolerance(o))
        ctx.close()


def test_rollout_with_resync(gws):
    """Per-step parity along a trajectory: every 10th step of 60 restart the device from the oracle state."""
    box = 0.4
    o = Oracle(box).setup_scene()
    for _ in range(6):
        ctx = make_ctx(gws, box, o.pos, o.vel)
        o.step(1)
        ctx.step(1)
        rec = ctx.download()
        assert np.abs(rec["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
        o.step(9)
        ctx.close()


def test_fused_step_equals_phase_path_bitwise(gws):
    o = state_after(0.4, 5)
    a = make_ctx(gws, 0.4, o.pos, o.vel)
    b = make_ctx(gws, 0.4, o.pos, o.vel)
    for _ in range(4):
        a.update_grid(); a.density_pressure(); a.forces(); a.collisions(); a.integrate()
    b.step(1)          # direct launches
    b.step(3)          # CUDA graph
    ra, rb = a.download(), b.download()
    for f in ("position", "velocity", "acceleration", "density", "pressure", "cell_id"):
        assert np.array_equal(ra[f].view(np.uint32), rb[f].view(np.uint32)), f
    assert b.counter("graph_launches") == 3


def test_determinism_and_input_order_independence(gws):
    """The canonical (cell,id) order makes results a pure function of the state: shuffling the upload
    order or re-running gives bit-identical outputs."""
    o = state_after(0.4, 7)
    pos, vel = o.pos, o.vel
    n = len(pos)
    a = make_ctx(gws, 0.4, pos, vel)
    rng = np.random.default_rng(0xC0FFEE)
    perm = rng.permutation(n)
    b = gws.SphContext(0.4, n)
    b.upload(gws.particles_from_arrays(pos[perm], vel[perm], ids=perm))
    a.step(5); b.step(5)
    ra, rb = a.download(), b.download()
    assert np.array_equal(ra["position"].view(np.uint32), rb["position"].view(np.uint32))
    assert np.array_equal(ra["density"].view(np.uint32), rb["density"].view(np.uint32))


@pytest.mark.parametrize("steps", [0, 1, 10, 100])
def test_brute_force_config(gws, steps):
    """config 2: all-pairs semantics at 32.5K particles; neighbour sets equal the grid walk and the oracle on the
    states SURVEY section 8(d) names (steps 0, 1, 10, 100)."""
    box = 1.14
    o = state_after(box, steps)
    assert o.n == 32500
    pos, vel = o.pos, o.vel
    ctx = make_ctx(gws, box, pos, vel)
    ctx.brute_density_pressure()
    brute_counts = ctx.brute_neighbour_counts()
    rho_b, prs_b, _ = ctx.density_pressure_accel()
    ctx.brute_forces()
    _, _, acc_b = ctx.density_pressure_accel()
    ctx.integrate()
    rec_b = ctx.download()
    ctx.upload(gws.particles_from_arrays(pos, vel))
    ctx.update_grid(); ctx.density_pressure()
    grid_counts, _ = ctx.neighbours(lists=False)
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    oc, ol = o.neighbours()
    assert np.array_equal(brute_counts, grid_counts)
    assert np.array_equal(brute_counts, oc)
    assert_neighbour_sets(ctx, oc, ol)  # the grid path's sets (production hit words and the validation walk) on this state
    check_density(o, rho_b, prs_b)
    check_acc(o.acc_sph, o.acc_scale, acc_b)
    o.integrate()
    assert np.abs(rec_b["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457


def test_fountain_through_simulator(gws):
    """Fountain scene driven through CCUDAParticleSimulator (emission on the host, append on the device)."""
    box = 0.4
    sim = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    o = Oracle(box, FOUNTAIN).setup_scene()
    steps = 40
    sim.step(steps)
    o.step(steps)
    sim.sync_host()
    hp = sim.host_particles()
    assert sim.n == o.n == 7 * steps
    assert np.array_equal(hp["id"], np.arange(sim.n, dtype=np.uint32))
    # the newest batch has only been integrated once: tight; older particles have been through up to 40
    # steps of a chaotic system (accelerations of 1e3..1e4 m/s^2 at the nozzle), so compare statistically
    err = np.abs(hp["position"][:, :3] - o.pos).max(axis=1)
    assert err[-7:].max() <= 1e-5
    assert np.percentile(err, 95) <= 0.0457 and np.isfinite(hp["position"]).all()
    assert np.abs(hp["position"][:, :3].mean(axis=0) - o.pos.mean(axis=0)).max() <= 0.1 * 0.0457
    ke_g = 0.5 * 0.02 * (hp["velocity"][:, :3].astype(np.float64) ** 2).sum()
    assert abs(ke_g / o.stats()["ke"] - 1) <= 0.05


def test_fountain_step_many_emits_on_the_device_bitwise(gws):
    """stepMany on a filling fountain: device-side emission (sph_set_emitter / k_emit) + the fused step, no per-step
    upload and no phase path — bit for bit what step() (host emission + append + phases) produces, through the cap."""
    box = 0.3   # max 784 particles: the emitter runs dry after 111 steps
    a = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    b = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    a.step(130)
    b.step_many(60)
    b.step_many(70)
    assert a.n == b.n == 7 * 111 and a.iteration == b.iteration == 130
    assert b.context().counter("graph_launches") > 0        # once the count is stable the graph path takes over
    a.sync_host(); b.sync_host()
    ha, hb = a.host_particles(), b.host_particles()
    for f in ("position", "velocity", "acceleration", "density", "pressure", "id", "cell_id"):
        assert np.array_equal(ha[f].view(np.uint32), hb[f].view(np.uint32)), f
    # mixing the two entry points keeps host and device counts in lock-step
    c = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    c.step(3); c.step_many(5); c.step(2); c.step_many(120)
    c.sync_host()
    assert np.array_equal(c.host_particles()["position"].view(np.uint32), ha["position"].view(np.uint32))


def test_fountain_config3_at_size_against_oracle(gws):
    """BASELINE configs[3]: fountain, box 2.28 (max 250 000 particles), 64 nozzles (emission multiplier), run on the
    device until more than 100 000 particles are in flight, then ONE step of both implementations from that state:
    keys, cell ranges, permutation, neighbour sets (production mask) exact; density, pressure, acceleration within
    rel 1e-5; positions within the acceleration tolerance."""
    box = 2.28
    sim = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    sim.set_emission_multiplier(64)
    assert sim.max_count == 250000
    sim.step_many(230)                                   # 64 * 7 * 230 = 103 040 particles, no host transfer
    assert sim.n == 64 * 7 * 230
    sim.sync_host()
    hp = sim.host_particles()
    assert np.array_equal(hp["id"], np.arange(sim.n, dtype=np.uint32)) and np.isfinite(hp["position"]).all()
    pos, vel = hp["position"][:, :3].copy(), hp["velocity"][:, :3].copy()
    o = Oracle(box, FOUNTAIN).set_state(pos, vel)
    ctx = make_ctx(gws, box, pos, vel, cap=sim.n)
    o.update_grid(); ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    cs, perm = o.cells()
    assert np.array_equal(ctx.cell_start(), cs) and np.array_equal(ctx.permutation().astype(np.int32), perm)
    o.update_density_pressure(); ctx.density_pressure()
    oc, ol = o.neighbours()
    assert_neighbour_sets(ctx, oc, ol)
    rho, prs, _ = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    o.update_forces(); ctx.forces()
    check_acc(o.acc_sph, o.acc_scale, ctx.density_pressure_accel()[2])
    ctx.integrate()
    check_acc(o.acc, o.acc_scale, ctx.density_pressure_accel()[2])
    rec = ctx.download()
    tol = pos_tolerance(o)
    o.integrate()
    assert np.all(np.abs(rec["position"][:, :3] - o.pos) <= tol)
    # and the next emission continues the ids on the device
    sim.step_many(1)
    assert sim.n == 64 * 7 * 231


def test_step_many_keeps_the_mirror_modes(gws):
    """stepMany() in a non-resident mirror mode refreshes the host mirror after the batch, and in
    RoundTrip the (canonical) host vector goes up first — a following step() must not rewind the simulation."""
    box = 0.4
    o = Oracle(box).setup_scene()
    sim = gws.Simulator("cuda", box).setup_scene()
    sim.set_mirror_mode(2)
    sim.step_many(6); o.step(6)
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    sim.step(2); o.step(2)                               # uploads the mirror: it must be the state after 6 steps
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    sim.set_mirror_mode(1)
    sim.step_many(4); o.step(4)
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 2e-4 * 0.0457
    assert sim.iteration == 12


def test_simulator_phase_path_and_mirror_modes(gws):
    box = 0.4
    o = Oracle(box).setup_scene()
    sim = gws.Simulator("cuda", box).setup_scene()
    assert "B200" in sim.device or "NVIDIA" in sim.device
    sim.set_profiling(True, 1)
    sim.set_mirror_mode(2)  # upload + download every step (the reference OpenCL path's semantics)
    sim.step(3)
    o.step(3)
    hp = sim.host_particles()
    assert np.abs(hp["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    ev = sim.events()
    assert ev.shape == (3, 7) and np.all(ev[:, 2:5] > 0) and np.all(ev[:, 5] == 0)
    # decimated viewer refresh: download mode with stride 4 only touches the mirror on every 4th step
    sim.set_mirror_mode(1)
    sim.set_mirror_stride(4)
    before = sim.host_particles()["position"].copy()
    sim.step(1)           # iteration 4 -> refresh
    o.step(1)
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    before = sim.host_particles()["position"].copy()
    sim.step(2)           # iterations 5, 6 -> mirror untouched
    o.step(2)
    assert np.array_equal(before, sim.host_particles()["position"])
    sim.set_mirror_stride(1)
    # resident fused path continues from the same state
    sim.set_mirror_mode(0)
    sim.step_many(4)
    o.step(4)
    sim.sync_host()
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 2e-4 * 0.0457
    assert sim.iteration == 10


def test_gravity_vector(gws):
    o = state_after(0.4, 2)
    ctx = make_ctx(gws, 0.4, o.pos, o.vel)
    g = (-1.0, -9.80665, 0.5)
    o.set_gravity(g); ctx.set_gravity(g)
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    ctx.update_grid(); ctx.density_pressure(); ctx.forces()
    check_acc(o.acc_sph, o.acc_scale, ctx.density_pressure_accel()[2])


def test_edge_cases(gws):
    # empty state, single particle, particles outside the box (clamped keys), capacity and call-order errors
    ctx = gws.SphContext(0.4, 16)
    ctx.upload(gws.particles_from_arrays(np.zeros((0, 3))))
    assert ctx.step(2) == 0.0 and ctx.n == 0
    with pytest.raises(gws.SphError):
        ctx.upload(gws.particles_from_arrays(np.zeros((17, 3))))
    pos = np.array([[0, 0, 0], [5.0, -5.0, 0.1], [-0.2, -0.2, -0.2], [0.19999, 0.2, 0.25]], dtype=np.float32)
    ctx.upload(gws.particles_from_arrays(pos))
    with pytest.raises(gws.SphError):
        ctx.forces()
    o = Oracle(0.4).set_state(pos)
    o.update_grid(); ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    o.update_density_pressure(); ctx.density_pressure()
    rho, prs, _ = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    assert abs(rho[0] - 328.2934) < 1e-3
    ctx.append(gws.particles_from_arrays(np.array([[0.01, 0.0, 0.0]], dtype=np.float32), ids=[4]))
    assert ctx.n == 5
    ctx.step(1)
    assert np.array_equal(np.sort(ctx.permutation()), np.arange(5))


@pytest.mark.parametrize("variant,n_clump", [(0, 400), (1, 400), (1, 1500)])
def test_dense_cell_overflow_path(gws, variant, n_clump):
    """Very dense clump: more hits than the variant-0 queue holds (400) and more candidates than the
    1024-slot hit bitmask of the production kernels holds (1500) -> their overflow paths."""
    rng = np.random.default_rng(1234)
    pos = (rng.random((n_clump, 3), dtype=np.float32) - 0.5) * np.float32(0.03)
    o = Oracle(0.4).set_state(pos)
    ctx = make_ctx(gws, 0.4, pos, np.zeros_like(pos), variant=variant)
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    ctx.update_grid(); ctx.density_pressure(); ctx.forces()
    oc, ol = o.neighbours()
    assert oc.max() > 100
    assert_neighbour_sets(ctx, oc, ol, variant)
    if variant == 1:
        assert (ctx.counter("overflow_particles") > 0) == (n_clump > 1000)  # 1500: the overflow list is what gets decoded
    rho, prs, acc = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    check_acc(o.acc_sph, o.acc_scale, acc)


def test_full_size_properties_1m(gws):
    """BASELINE config 3 (1,011,240 particles): size-independent properties instead of an oracle run."""
    box = 3.62
    sim = gws.Simulator("cuda", box).setup_scene()
    n = sim.n
    assert n == 1011240
    ctx = sim.context()
    sim.step_many(20)
    ctx.update_grid()
    keys, perm, cs = ctx.keys(), ctx.permutation(), ctx.cell_start()
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))          # a permutation
    assert np.array_equal(np.diff(cs), np.bincount(keys, minlength=ctx.n_cells))  # ranges == histogram
    assert cs[0] == 0 and cs[-1] == n
    sk = keys[perm]
    assert np.all(np.diff(sk) >= 0)                                               # sorted by cell
    same = np.diff(sk) == 0
    assert np.all(np.diff(perm.astype(np.int64))[same] > 0)                       # ids ascend inside a cell
    ctx.density_pressure()
    rho, _, _ = ctx.density_pressure_accel()
    counts, _ = ctx.neighbours(lists=False)
    assert rho.min() >= 328.29 and counts.min() >= 1 and np.isfinite(rho).all()
    st = ctx.stats()
    assert np.isfinite(st["ke"]) and st["fill"] <= box * 1.01
    # keys agree with the oracle's key formula on the downloaded positions (fp64 math on the host)
    ctx.forces(); ctx.integrate()
    rec = ctx.download()
    assert np.isfinite(rec["position"]).all()


def test_cell_ranges_on_a_long_grid_with_many_scan_tiles(gws):
    """K2 (k_scan) looks back over 512 predecessor tiles per round; a 1.6 x 0.8 x 160 tank has ~9 M bins = more than
    1 000 tiles, i.e. several rounds.  cell ranges == histogram of the keys, step after step (a race in the look-back
    shows up as a wrong range somewhere in a few launches)."""
    sim = gws.Simulator("cuda", (1.6, 0.8, 160.0)).setup_scene()
    n = sim.n
    ctx = sim.context()
    assert 4 * ctx.n_cells > 1000 * 8192
    for rep in range(12):
        sim.step_many(3)
        ctx.update_grid()
        keys, cs = ctx.keys(), ctx.cell_start()
        assert cs[0] == 0 and cs[-1] == n, rep
        assert np.array_equal(np.diff(cs), np.bincount(keys, minlength=ctx.n_cells)), rep
    perm = ctx.permutation()
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    assert np.all(np.diff(keys[perm]) >= 0)


def test_rollout_statistics_1000_steps(gws):
    """BASELINE north_star: long rollouts are chaotic, so 1000 steps are compared statistically.  Definitions of
    SURVEY.md §8c, sampled every 10 steps (eventLoggerStride): kinetic energy 0.5 m sum|v|^2, centre of mass,
    fill height max(y)+b/2 and its robust form, the 95th percentile of y+b/2.  Stated tolerances: COM within 0.25 h per
    axis at every sample, fill height (a max over particles, i.e. one splashing particle) within 2 h, its 95th
    percentile within 0.25 h, KE within 2 % of the run's peak KE at every sample and
    within 5 % relative during the collapse (first 50 steps; afterwards KE decays by five orders of magnitude and
    the two chaotic trajectories only agree in the absolute sense).  Measured on B200: 0.05 h, 1.1 h, 0.6 %, 2 %."""
    box, h = 0.4, 0.0457
    o = Oracle(box).setup_scene()
    sim = gws.Simulator("cuda", box).setup_scene()
    ctx = sim.context()
    ke_o, ke_g, com_err, fill_err, fill95_err = [], [], [], [], []
    for _ in range(100):
        o.step(10)
        sim.step_many(10)
        so, sg = o.stats(), ctx.stats()
        ke_o.append(so["ke"]); ke_g.append(sg["ke"])
        com_err.append(np.abs(so["com"] - sg["com"]).max())
        fill_err.append(abs(so["fill"] - sg["fill"]))
        fill95_err.append(abs(so["fill95"] - ctx.fill_height_percentile(0.95)))
    ke_o, ke_g = np.array(ke_o), np.array(ke_g)
    assert max(com_err) <= 0.25 * h, max(com_err)
    assert max(fill_err) <= 2.0 * h, max(fill_err)
    assert max(fill95_err) <= 0.25 * h, max(fill95_err)
    assert np.abs(ke_g - ke_o).max() <= 0.02 * ke_o.max()
    assert np.all(np.abs(ke_g[:5] / ke_o[:5] - 1) <= 0.05)
    assert np.isfinite(ke_g).all() and sim.iteration == 1000


def test_fountain_at_scale_with_nozzle_array(gws):
    """BASELINE configs[3]: fountain with continuous emission and the six walls, scaled up with the nozzle array
    (emission multiplier extension): 64 nozzles x 7 particles per step until the scene is full."""
    box = 1.14  # max 32 500 particles
    sim = gws.Simulator("cuda", box, scenario=gws.FOUNTAIN).setup_scene()
    sim.set_emission_multiplier(64)
    assert sim.n == 0 and sim.max_count == 32500
    sim.step(100)
    assert sim.n == min(64 * 7 * 100, (32500 - 7) // 7 * 7 + 7) or sim.n <= 32500
    sim.sync_host()
    hp = sim.host_particles()
    assert np.isfinite(hp["position"]).all() and np.isfinite(hp["velocity"]).all()
    # the walls keep the water in the box (penalty walls are soft: the jet hits the lid at 3.6 m/s; allow 3 h)
    assert np.abs(hp["position"][:, :3]).max() <= box / 2 + 3 * 0.0457
    assert np.array_equal(np.sort(hp["id"]), np.arange(sim.n, dtype=np.uint32))
    ctx = sim.context()
    ctx.update_grid(); ctx.density_pressure()
    counts, _ = ctx.neighbours(lists=False)
    assert counts.min() >= 1 and hp["density"].max() > 328.0


@pytest.mark.parametrize("seed,n,box", [(0xC0FFEE, 6000, 0.5), (7, 2500, (0.6, 0.25, 0.4)), (99, 4000, 0.4)])
def test_random_states_with_awkward_particles(gws, seed, n, box):
    """Seeded random clouds instead of lattice-born states: non-uniform density, particles outside the box
    (clamped into edge cells like the reference does), particles exactly on cell faces and box walls, duplicated
    x/y/z coordinates, and non-contiguous ids in shuffled upload order."""
    rng = np.random.default_rng(seed)
    b = np.array([box] * 3 if np.isscalar(box) else box, dtype=np.float32)
    pos = ((rng.random((n, 3), dtype=np.float32) - 0.5) * b * np.float32(0.9)).astype(np.float32)
    pos[: n // 4] *= np.float32(0.35)                                   # a dense core (60+ neighbours)
    h = np.float32(0.0457)
    k = n // 10
    # x and z snapped to multiples of h (cell faces when b/2 is a multiple of h, generic otherwise); y stays random so
    # that no two particles coincide (r = 0 gives 0/0 = NaN in the reference as well)
    pos[n // 2:n // 2 + k, 0::2] = (np.floor(pos[n // 2:n // 2 + k, 0::2] / h) * h).astype(np.float32)
    pos[-40:-20] = (pos[-40:-20] + np.sign(pos[-40:-20]) * b).astype(np.float32)               # outside the box
    pos[-20:-10, 0] = -b[0] / 2                                          # exactly on the left wall
    pos[-10:, 1] = pos[-11, 1]                                           # shared y coordinate
    vel = ((rng.random((n, 3), dtype=np.float32) - 0.5) * np.float32(2.0)).astype(np.float32)
    o = Oracle(tuple(float(v) for v in b)).set_state(pos, vel)
    ids = rng.permutation(n)                                             # upload order != id order
    ctx = gws.SphContext(tuple(float(v) for v in b), n)
    ctx.upload(gws.particles_from_arrays(pos[ids], vel[ids], ids=ids))
    o.update_grid(); ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    cs, perm = o.cells()
    assert np.array_equal(ctx.cell_start(), cs) and np.array_equal(ctx.permutation().astype(np.int32), perm)
    o.update_density_pressure(); ctx.density_pressure()
    oc, ol = o.neighbours()
    assert_neighbour_sets(ctx, oc, ol)
    rho, prs, _ = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    o.update_forces(); ctx.forces()
    assert np.isfinite(o.acc).all()
    check_acc(o.acc_sph, o.acc_scale, ctx.density_pressure_accel()[2])
    ctx.integrate()
    acc_tot = ctx.density_pressure_accel()[2]
    check_acc(o.acc, o.acc_scale, acc_tot)
    rec = ctx.download()
    tol = pos_tolerance(o)
    o.integrate()
    assert np.all(np.abs(rec["position"][:, :3] - o.pos) <= tol)


def test_headless_cpp_bench_runs(gws):
    """core/sph_bench (the headless C++ driver of the reference-facing simulator) on the small dam break."""
    import json
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gmu-water-simulation_b200", "sph_bench")
    out = subprocess.run([exe, "--box", "0.9", "--steps", "50", "--warmup", "5"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["particles"] == 16000 and line["particle_steps_per_s"] > 1e6
    out = subprocess.run([exe, "--box", "0.4", "--scenario", "fountain", "--steps", "30", "--warmup", "0", "--phases"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["particles"] == 210 and line["phase_ms"]["density"] > 0 and line["phase_ms"]["collisions"] == 0


@pytest.mark.parametrize("variant", [1, 0])
def test_non_finite_particle_stays_confined(gws, variant):
    """A particle with a NaN coordinate is nobody's neighbour in the reference (comparisons with NaN are false) and
    lands in cell 0 after clamping; the rest of the scene must be unaffected - same sets, same densities."""
    o = state_after(0.4, 6)
    pos, vel = o.pos.copy(), o.vel.copy()
    bad = 777
    pos[bad] = [np.nan, 0.01, 0.02]
    o = Oracle(0.4).set_state(pos, vel)
    ctx = make_ctx(gws, 0.4, pos, vel, variant=variant)
    o.update_grid(); ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    o.update_density_pressure(); ctx.density_pressure()
    oc, _ = o.neighbours(lists=False)
    gc, _ = ctx.neighbours(lists=False)
    others = np.arange(o.n) != bad
    assert np.array_equal(gc[others], oc[others]) and oc[bad] == 0
    rho, prs, _ = ctx.density_pressure_accel()
    assert np.all(np.abs(rho[others] - o.density[others]) <= RTOL * o.density[others])
    o.update_forces(); ctx.forces(); ctx.integrate()
    rec = ctx.download()
    assert np.isfinite(rec["position"][others, :3]).all() and np.isfinite(rec["acceleration"][others, :3]).all()
    # the reference leaves such a particle non-finite for good (NaN + anything = NaN): so do the phase path and the
    # FUSED whole-step path (sph_step), bit for bit alike for every other particle
    assert np.isnan(rec["position"][bad, 0])
    fused = make_ctx(gws, 0.4, pos, vel, variant=variant)
    fused.step(1)
    rf = fused.download()
    assert np.isnan(rf["position"][bad, 0]) and np.isfinite(rf["position"][others, :3]).all()
    for f in ("position", "velocity", "acceleration", "density"):
        assert np.array_equal(rec[f][others].view(np.uint32), rf[f][others].view(np.uint32)), f
    fused.step(3)
    rf = fused.download()
    assert np.isnan(rf["position"][bad, 0]) and np.isfinite(rf["position"][others, :3]).all()


def test_exact_percentile_of_fill_height(gws):
    """sph_fill_height_percentile: the element of rank floor(q (n - 1)) of y + b/2 (oracle_stats' definition)."""
    o = state_after(0.9, 40)
    ctx = make_ctx(gws, 0.9, o.pos, o.vel)
    y = np.sort(o.pos[:, 1].astype(np.float64)) + np.float32(0.9) / 2.0
    for q in (0.0, 0.05, 0.5, 0.95, 1.0):
        want = y[int(q * (o.n - 1))]
        assert abs(ctx.fill_height_percentile(q) - want) <= 1e-6, q
    assert abs(ctx.fill_height_percentile(0.95) - o.stats()["fill95"]) <= 1e-6


def test_particle_id_precondition_is_checked(gws):
    """by-id read-backs index host arrays with the particle ids.  An id >= max_particles must never be
    written through; the next by-id read-back reports it.  Ids in [n, max_particles) are skipped, not written."""
    o = state_after(0.4, 3)
    n = o.n
    ctx = gws.SphContext(0.4, n + 8)
    ids = np.arange(n, dtype=np.uint32)
    ids[5] = 4_000_000_000
    ctx.upload(gws.particles_from_arrays(o.pos, o.vel, ids=ids))
    ctx.update_grid(); ctx.density_pressure()
    with pytest.raises(gws.SphError, match="max_particles"):
        ctx.keys()
    with pytest.raises(gws.SphError, match="max_particles"):
        ctx.neighbours(lists=False)
    with pytest.raises(gws.SphError, match="max_particles"):
        ctx.download()
    ids[5] = n + 3                                        # inside the capacity but not a permutation of 0..n-1
    ctx.upload(gws.particles_from_arrays(o.pos, o.vel, ids=ids))   # a full upload re-arms the check
    ctx.update_grid(); ctx.density_pressure()
    keys = ctx.keys()
    ref = make_ctx(gws, 0.4, o.pos, o.vel)
    ref.update_grid()
    good = np.arange(n) != 5
    assert np.array_equal(keys[good], ref.keys()[good])
    ctx.step(2)                                           # the simulation itself does not depend on the id values


def test_full_size_parity_against_oracle_1m(gws):
    """The headline workload itself: 1,011,240 particles, 200 steps into the dam break (the benchmarked state).
    One more step on both sides from the same state: keys, cell ranges, neighbour counts exact; density, pressure,
    acceleration within rel 1e-5; positions within the acceleration tolerance.  (~15 s of oracle time.)"""
    box = 3.62
    sim = gws.Simulator("cuda", box).setup_scene()
    sim.step_many(200)
    sim.sync_host()
    hp = sim.host_particles()
    assert np.array_equal(hp["id"], np.arange(sim.n, dtype=np.uint32))
    pos, vel = hp["position"][:, :3].copy(), hp["velocity"][:, :3].copy()
    o = Oracle(box).set_state(pos, vel)
    ctx = sim.context()
    o.update_grid(); ctx.update_grid()
    assert np.array_equal(ctx.keys(), o.keys())
    cs, perm = o.cells()
    assert np.array_equal(ctx.cell_start(), cs)
    assert np.array_equal(ctx.permutation().astype(np.int32), perm)
    o.update_density_pressure(); ctx.density_pressure()
    oc, ol = o.neighbours()                     # ~6.7e7 ids
    gc, _ = ctx.neighbours(lists=False)
    assert np.array_equal(gc, oc) and oc.mean() > 30
    mc, ml = ctx.neighbours(source="mask")      # the production hit words at full size, as sets
    assert np.array_equal(mc, oc) and np.array_equal(ml, ol)
    del ol, ml
    rho, prs, _ = ctx.density_pressure_accel()
    check_density(o, rho, prs)
    o.update_forces(); ctx.forces()
    check_acc(o.acc_sph, o.acc_scale, ctx.density_pressure_accel()[2])
    ctx.integrate()
    rec = ctx.download()
    tol = pos_tolerance(o)
    o.integrate()
    assert np.all(np.abs(rec["position"][:, :3] - o.pos) <= tol)
    assert ctx.counter("overflow_particles") == 0


def test_async_download_is_a_consistent_snapshot(gws):
    """Viewer bridge (row f2): the overlapped read-back delivers exactly the state at the moment it was requested,
    while later steps are already running; a second request queues behind the first on the device."""
    box = 0.9
    o = state_after(box, 3)
    ctx = make_ctx(gws, box, o.pos, o.vel)
    ctx.step(2)
    ref = ctx.download().copy()                      # blocking read-back of the same moment
    out = np.zeros(o.n, dtype=gws.PARTICLE_DTYPE)
    ctx.pin(out)
    ctx.download_async(out)
    ctx.step(5, timed=False)                         # keeps the device busy while the copy is in flight
    assert ctx.download_wait() == o.n
    assert out.tobytes() == ref.tobytes()
    ref2 = None
    ctx.download_async(out)                          # state after 7 steps
    ctx.step(1, timed=False)
    ctx.download_async(out)                          # state after 8 steps: waits for the first copy on the device
    ctx.download_wait()
    ref2 = ctx.download()
    assert out.tobytes() == ref2.tobytes()
    assert ctx.download_wait() == o.n                # nothing in flight: returns immediately
    ctx.close()


def test_simulator_async_mirror_mode(gws):
    box = 0.4
    o = Oracle(box).setup_scene()
    sim = gws.Simulator("cuda", box).setup_scene()
    sim.set_mirror_mode(3)                           # AsyncDownload
    sim.set_mirror_stride(2)
    sim.step(4); o.step(4)                           # refreshes requested after steps 2 and 4
    hp = sim.host_particles()                        # completes the copy in flight
    assert np.abs(hp["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    sim.step(1); o.step(1)                           # step 5: no refresh, the mirror still shows step 4
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() > 0
    sim.sync_host()                                  # blocking: current state
    assert np.abs(sim.host_particles()["position"][:, :3] - o.pos).max() <= 1e-4 * 0.0457
    sim.close()


def test_round_trip_mode_continues_from_the_device_state(gws):
    """Switching to RoundTrip after resident steps must not rewind the simulation to the stale host mirror."""
    box = 0.4
    o = Oracle(box).setup_scene()
    sim = gws.Simulator("cuda", box).setup_scene()
    sim.step_many(12)                                # resident: the host mirror still holds the initial lattice
    sim.set_mirror_mode(2)                           # brings the mirror up to date before it becomes canonical
    sim.step(3)
    o.step(15)
    hp = sim.host_particles()
    # a rewind would show up as centimetres (12 steps of free fall); chaos after 15 steps is micrometres
    assert np.abs(hp["position"][:, :3] - o.pos).max() <= 1e-3
    assert sim.iteration == 15
    sim.close()
