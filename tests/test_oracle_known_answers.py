"""The oracle against every known answer derivable from the reference source.

The reference has no tests or golden vectors of its own ("parity unpinned"), so these known answers,
the committed fixtures in tests/golden/ and the independent numpy restatement are what pins the oracle.
"""
import json
import os

import numpy as np
import pytest

from oracle_binding import DAM_BREAK, FOUNTAIN, Oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize(
    "box,count,res",
    [(0.1, 50, 3), (0.2, 243, 5), (0.3, 784, 7), (0.4, 1620, 9), (0.5, 2904, 11), (0.6, 5103, 14),
     (0.7, 7688, 16), (0.8, 11664, 18), (0.9, 16000, 20), (1.0, 21296, 22)],
)
def test_particle_counts_and_grid(box, count, res):
    # src/CBaseParticleSimulator.cpp:41 (count formula + assert :56), :27-31 (grid); 1620/16000 are in doc.pdf pp.6-7
    o = Oracle(box).setup_scene()
    assert o.n == count
    assert o.max_count == count
    assert o.grid_res == (res, res, res)


def test_bench_scene_sizes():
    assert Oracle(1.14).setup_scene().n == 32500
    o = Oracle((4.56, 4.56, 9.13), FOUNTAIN)
    assert o.grid_res == (100, 100, 200)


def test_constants():
    o = Oracle(0.4)
    poly6, spiky, visc, h2 = o.constants()
    assert h2.view(np.uint32) == 0x3B08DF0C  # fl32(0.0457f*0.0457f) == the OpenCL literal 0.00208849f
    assert np.float32(poly6) == np.float32(1801917956096.0)
    assert np.float32(spiky) == np.float32(-1572408960.0)
    assert visc == -spiky
    assert abs(poly6 / 1.8019179624e12 - 1) < 1e-9


def test_all_particles_start_in_cell_zero():
    # src/CBaseParticleSimulator.cpp:69-72
    o = Oracle(0.4).setup_scene()
    cs, ids = o.cells()
    assert cs[1] == o.n and np.array_equal(ids, np.arange(o.n))


def test_isolated_particle():
    o = Oracle(0.4)
    o.add_particle(0.0, 0.0, 0.0)
    o.update_grid()
    o.update_density_pressure()
    o.update_forces()
    assert abs(float(o.density[0]) - 328.2934) < 1e-3  # m * poly6 * h^6
    assert abs(float(o.pressure[0]) - (-2009.99)) < 1e-2
    a = o.acc[0]
    assert a[0] == 0 and a[2] == 0 and abs(a[1] - (-9.80665)) < 2e-6


def test_two_particle_closed_form():
    """Pair of particles: density, pressure and both force terms against the formulas of the reference evaluated in
    fp64 (src/CCPUParticleSimulator.cpp:9-30 kernels, :122-134 density/pressure, :174-195 forces).  Also pins the
    sign conventions: a compressed pair (p > 0 needs rho > rho0; here p < 0) attracts, viscosity pulls velocities together."""
    import math
    h, m, k, rho0, mu, g = float(np.float32(0.0457)), 0.02, 3.0, float(np.float32(998.29)), 3.5, -9.80665
    poly6 = 315.0 / (64.0 * math.pi * h ** 9)
    spiky = -45.0 / (math.pi * h ** 6)
    visc = 45.0 / (math.pi * h ** 6)
    xi, xj = np.array([0.0, 0.0, 0.0]), np.array([0.02, 0.01, -0.005])
    vi, vj = np.array([0.3, -0.2, 0.1]), np.array([-0.1, 0.4, 0.25])
    o = Oracle(0.9)
    o.set_state(np.array([xi, xj], dtype=np.float32), np.array([vi, vj], dtype=np.float32))
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    xi, xj = np.float32(xi).astype(np.float64), np.float32(xj).astype(np.float64)
    vi, vj = np.float32(vi).astype(np.float64), np.float32(vj).astype(np.float64)
    d = xi - xj
    r2 = float(d @ d)
    r = math.sqrt(r2)
    assert r < h
    rho = m * poly6 * ((h * h) ** 3 + (h * h - r2) ** 3)     # self + the one neighbour, identical for both
    p = k * (rho - rho0)
    assert np.allclose(o.density, rho, rtol=2e-6) and np.allclose(o.pressure, p, rtol=1e-5)
    grad = spiky * (h - r) ** 2 * d / r                        # WspikyGradient(x_i - x_j)
    f_p = (p / rho ** 2 + p / rho ** 2) * grad
    f_v = visc * (h - r) * (vj - vi) / rho                     # WviscosityLaplacian * (v_j - v_i) / rho_j
    a_i = ((-m * rho) * f_p + (mu * m) * f_v + np.array([0.0, g, 0.0]) * rho) / rho
    a_j = ((-m * rho) * (-f_p) + (mu * m) * (-f_v) + np.array([0.0, g, 0.0]) * rho) / rho
    assert np.allclose(o.acc_sph[0], a_i, rtol=2e-5, atol=1e-6) and np.allclose(o.acc_sph[1], a_j, rtol=2e-5, atol=1e-6)
    assert p < 0 and np.dot(a_i - np.array([0.0, g, 0.0]) - (mu * m) * f_v / rho, xj - xi) > 0   # negative pressure attracts
    assert np.dot((mu * m) * f_v, vj - vi) > 0                                                  # viscosity damps the relative motion


def test_lattice_step0_neighbour_statistics():
    # probe values recorded in SURVEY.md §8c for N=1620 (box 0.4)
    o = Oracle(0.4).setup_scene()
    o.update_grid()
    o.update_density_pressure()
    counts, flat = o.neighbours()
    assert counts.sum() == 38608 == flat.size
    assert counts[0] == 11 and counts.min() == 9 and counts.max() == 33
    assert abs(counts.mean() - 23.83) < 0.01
    assert abs(float(o.density.max()) - 1692.763) < 2e-3
    assert abs(float(o.pressure.max()) - 2083.419) < 6e-3


def test_wall_force_at_rest():
    # particle at distance s < 0.01 inside the left wall, at rest: 10000 * (0.01 - s) in +x
    b = 0.4
    o = Oracle(b)
    s = 0.004
    o.add_particle(-np.float32(b) / 2 + np.float32(s), 0.0, 0.0)
    o.update_grid()
    o.update_density_pressure()
    o.update_forces()
    w = o.acc_wall[0]
    assert abs(w[0] - 10000.0 * (0.01 - s)) < 1e-2 and w[1] == 0 and w[2] == 0


def test_fountain_emission():
    # src/CBaseParticleSimulator.cpp:187-210
    o = Oracle(0.4, FOUNTAIN).setup_scene()
    assert o.n == 0 and o.max_count == 1620
    assert o.generate_particles() == 7
    p, v = o.pos, o.vel
    hp, h = np.float32(0.0457) / np.float32(2), np.float32(0.0457)
    expect = np.array([[0, 0], [-hp, 0], [hp, 0], [-h / 4, -hp], [h / 4, -hp], [-h / 4, hp], [h / 4, hp]], dtype=np.float32)
    assert np.array_equal(p[:, [0, 2]], expect)
    assert np.all(p[:, 1] == np.float32(-0.2))
    assert np.all(v[:, 1] == np.float32(0.4) * np.float32(3.2)) and np.all(v[:, [0, 2]] == 0)
    o.step(300)
    assert o.n == 1617  # stops once count >= max - 7


def test_grid_membership_matches_keys():
    o = Oracle(0.5).setup_scene()
    o.step(5)
    o.update_grid()
    keys = o.keys()
    cs, ids = o.cells()
    order = np.lexsort((np.arange(o.n), keys))
    assert np.array_equal(ids, order)
    assert np.array_equal(np.diff(cs), np.bincount(keys, minlength=o.n_cells))


def test_brute_equals_grid_on_small_scene():
    a = Oracle(0.3).setup_scene()
    a.step(3)
    b = Oracle(0.3).set_state(a.pos, a.vel)
    a.update_grid(); a.update_density_pressure(); a.update_forces()
    b.brute_density_pressure(); b.brute_forces()
    # same neighbour sets, different summation order
    np.testing.assert_allclose(b.density, a.density, rtol=2e-6)
    scale = np.maximum(np.linalg.norm(a.acc, axis=1), a.acc_scale)
    assert np.all(np.abs(b.acc - a.acc).max(axis=1) <= 1e-5 * scale)


def test_golden_fixtures():
    """Fixtures generated by tests/golden/make_golden.py from this oracle at commit time: guards the
    oracle against accidental drift (compiler flags, refactors)."""
    path = os.path.join(GOLDEN, "oracle_dam_break_0p3.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated")
    g = np.load(path)
    o = Oracle(0.3).setup_scene()
    for k in range(int(g["steps"])):
        o.step(1)
    assert np.array_equal(o.pos.view(np.uint32), g["pos"].view(np.uint32))
    assert np.array_equal(o.vel.view(np.uint32), g["vel"].view(np.uint32))
    assert np.array_equal(o.density.view(np.uint32), g["density"].view(np.uint32))


def test_threaded_variant_is_bit_identical():
    """The labelled all-cores variant of the oracle (NOT reference behaviour) must not change a single bit."""
    a = Oracle(0.4).setup_scene()
    b = Oracle(0.4).setup_scene()
    assert b.set_threads(4) == 4
    a.step(12)
    b.step(12)
    for x, y in ((a.pos, b.pos), (a.vel, b.vel), (a.acc, b.acc), (a.density, b.density)):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
