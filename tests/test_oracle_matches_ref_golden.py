"""The oracle restatement against vectors produced by the REFERENCE'S OWN code.

tests/golden/ref_*.npz were written by tests/golden/make_ref_golden.py from oracle/_ref/libsph_ref.so (the
unmodified sources of /root/reference compiled against oracle/qt_shim).  Unlike test_ref_pins_oracle.py this
needs neither /root/reference nor the built library, so it pins the oracle on every box: the oracle, run from
the scene start, must reproduce every stored snapshot bit for bit, including the order inside every grid cell.
"""
import os

import numpy as np
import pytest

from oracle_binding import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ref_dam_break_0p4", "ref_fountain_0p4", "ref_dam_break_0p9"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_vectors_bit_for_bit(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    o = Oracle(float(g["box"]), int(g["scenario"])).setup_scene()
    done = 0
    for k in [int(v) for v in g["steps"]]:
        o.step(k - 1 - done)
        assert o.n == int(g[f"s{k}_n_before"])
        assert np.array_equal(bits(o.pos), bits(g[f"s{k}_pos_before"])), f"{name} before step {k}: pos"
        assert np.array_equal(bits(o.vel), bits(g[f"s{k}_vel_before"])), f"{name} before step {k}: vel"
        o.step(1)
        done = k
        for field, val in (("pos", o.pos), ("vel", o.vel), ("acc", o.acc), ("density", o.density), ("pressure", o.pressure)):
            ref = g[f"s{k}_{field}"]
            assert np.array_equal(bits(val), bits(ref)), f"{name} step {k}: {field} differs in {(bits(val) != bits(ref)).sum()} words"
        cs, ids = o.cells_raw()
        assert np.array_equal(cs, g[f"s{k}_cell_start"]), f"{name} step {k}: cell_start"
        assert np.array_equal(ids, g[f"s{k}_ids"]), f"{name} step {k}: intra-cell order"
