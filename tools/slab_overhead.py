"""How much does the slab step loop itself cost?  One GPU, the tank of one rank's share of the 64M run
(4.56 x 4.56 x 18.28, 8,000,000 particles), stepped (a) by the plain graph path and (b) as a world-1 slab context
(same kernels, but the slab loop: per-step host reads of the layer starts, direct launches, boundary/interior split).

    python tools/slab_overhead.py [--steps 50] [--preroll 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--preroll", type=int, default=200)
ap.add_argument("--box", type=float, nargs=3, default=[4.56, 4.56, 146.23 / 8.0])
a = ap.parse_args()
out = {}
for mode in ("plain", "slab_world1", "slab_world1_sequential"):
    sim = gws.Simulator("cuda", tuple(a.box))
    if mode != "plain":
        sim.enable_slab(0, 1, bytes(128))
    sim.setup_scene()
    if mode == "slab_world1_sequential":
        sim.context().set_option("slab_overlap", 0)
    sim.step_many(a.preroll, timed=False)
    sim.step_many(5)
    ms = sim.step_many(a.steps)
    out[mode] = {"particles": sim.n, "ms_per_step": ms / a.steps}
    sim.close()
print(json.dumps(out))
