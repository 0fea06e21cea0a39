"""Pins the C++ oracle with a second, independently written restatement (tests/np_restatement.py): the two
must agree bit for bit, phase by phase, on small dam-break scenes (the reference itself cannot run here)."""
import numpy as np
import pytest

from np_restatement import NpSim
from oracle_binding import Oracle


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("box,steps", [(0.2, 8), (0.3, 3), ((0.25, 0.2, 0.3), 4)])
def test_bitwise_agreement(box, steps):
    o = Oracle(box).setup_scene()
    s = NpSim(box).setup_dam_break()
    assert o.n == len(s.pos) and tuple(s.res) == o.grid_res
    assert np.array_equal(bits(o.pos), bits(s.arrays()[0]))
    for k in range(steps):
        o.update_grid(); s.update_grid()
        cs, ids = o.cells()
        for (x, y, z), members in s.cells.items():
            c = x + y * s.res[0] + z * s.res[0] * s.res[1]
            assert sorted(members) == list(ids[cs[c]:cs[c + 1]]), f"cell membership differs at step {k}"
        o.update_density_pressure(); s.density_pressure()
        counts, flat = o.neighbours()
        off = np.concatenate([[0], np.cumsum(counts)])
        for i in range(o.n):
            assert list(flat[off[i]:off[i + 1]]) == s.nb[i], f"neighbour set of {i} differs at step {k}"
        o.update_forces(); s.forces()
        pos, vel, acc, rho, prs = s.arrays()
        assert np.array_equal(bits(o.density), bits(rho)), f"density differs at step {k}"
        assert np.array_equal(bits(o.pressure), bits(prs)), f"pressure differs at step {k}"
        assert np.array_equal(bits(o.acc), bits(acc)), f"acceleration differs at step {k}"
        o.integrate(); s.integrate()
        pos, vel, _, _, _ = s.arrays()
        assert np.array_equal(bits(o.pos), bits(pos)) and np.array_equal(bits(o.vel), bits(vel)), f"state differs after step {k}"
