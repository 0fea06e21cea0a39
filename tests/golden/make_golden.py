"""Generates the committed fixtures in tests/golden/ from the CPU oracle.

The reference has no golden vectors and cannot run here, so these fixtures pin the ORACLE against
drift (they are its own outputs at the time the oracle was validated against the known answers of
tests/test_oracle_known_answers.py).  Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_binding import Oracle  # noqa: E402


def main():
    steps = 20
    o = Oracle(0.3).setup_scene()
    o.step(steps)
    np.savez_compressed(os.path.join(HERE, "oracle_dam_break_0p3.npz"), steps=steps, pos=o.pos, vel=o.vel,
                        density=o.density, pressure=o.pressure, acc=o.acc)
    # per-phase snapshot at step 10 of box 0.4 for the GPU parity tests (inputs + expected outputs)
    o = Oracle(0.4).setup_scene()
    o.step(10)
    pos, vel = o.pos, o.vel
    o.update_grid(); o.update_density_pressure(); o.update_forces()
    cs, ids = o.cells()
    counts, _ = o.neighbours(lists=False)
    np.savez_compressed(os.path.join(HERE, "oracle_phase_0p4_step10.npz"), pos=pos, vel=vel, keys=o.keys(), cell_start=cs,
                        perm=ids, nb_counts=counts, density=o.density, pressure=o.pressure, acc_sph=o.acc_sph,
                        acc=o.acc, acc_scale=o.acc_scale)


if __name__ == "__main__":
    main()
