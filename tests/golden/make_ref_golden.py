"""Generates tests/golden/ref_*.npz from the REFERENCE'S OWN CPU simulator (oracle/_ref/libsph_ref.so: the
unmodified sources of /root/reference compiled against oracle/qt_shim, see oracle/Makefile target `ref`).

These are reference outputs, not oracle outputs: they travel to boxes where /root/reference does not exist and
pin (a) the oracle restatement bit for bit (tests/test_oracle_matches_ref_golden.py, CPU) and (b) the CUDA path
within the north-star tolerances (tests/test_gpu_parity.py::test_parity_against_reference_golden).

Each snapshot k stores the state BEFORE step k (pos_before, vel_before = the reference's state after step k-1)
and what the reference holds AFTER step k: density, pressure, acc (from step k's force phase), pos, vel, and the
grid as step k's updateGrid left it (cell_start, ids in the reference's own intra-cell order).

Run here (needs /root/reference):  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_binding  # noqa: E402

CASES = [
    # name, box, scenario, snapshot steps (1-based: snapshot k describes the k-th step)
    ("ref_dam_break_0p4", 0.4, ref_binding.DAM_BREAK, (1, 2, 10, 50, 120)),
    ("ref_fountain_0p4", 0.4, ref_binding.FOUNTAIN, (1, 50, 150)),
    ("ref_dam_break_0p9", 0.9, ref_binding.DAM_BREAK, (100,)),   # BASELINE configs[0], 16 000 particles
]


def main():
    for name, box, scenario, snaps in CASES:
        r = ref_binding.Reference(box, scenario).setup_scene()
        out = {"box": np.float32(box), "scenario": np.int32(scenario), "steps": np.array(snaps, dtype=np.int32)}
        done = 0
        for k in snaps:
            r.step(k - 1 - done)
            pos_b, vel_b = r.pos, r.vel
            r.step(1)
            done = k
            cs, ids = r.cells()
            n_before = len(pos_b)
            for field, val in (("pos_before", pos_b), ("vel_before", vel_b), ("pos", r.pos), ("vel", r.vel),
                               ("acc", r.acc), ("density", r.density), ("pressure", r.pressure),
                               ("cell_start", cs), ("ids", ids), ("n_before", np.int64(n_before))):
                out[f"s{k}_{field}"] = val
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
