"""Short driver for ncu: 1M dam break, pre-roll, then a few phase-by-phase steps (one launch per kernel)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--box", type=float, default=3.62)
ap.add_argument("--preroll", type=int, default=200)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--neighbour-variant", type=int, default=1)
ap.add_argument("--tuning", type=int, default=0)
a = ap.parse_args()
sim = gws.Simulator("cuda", a.box).setup_scene()
ctx = sim.context()
ctx.set_option("use_graph", 0)
ctx.set_option("neighbour_variant", a.neighbour_variant)
ctx.set_option("tuning", a.tuning)
for _ in range(a.preroll):
    ctx.step(1, timed=False)
ctx.synchronize()
ms = {"grid": 0.0, "density": 0.0, "forces": 0.0, "integrate": 0.0}
for _ in range(a.steps):
    ms["grid"] += ctx.update_grid()
    ms["density"] += ctx.density_pressure()
    ms["forces"] += ctx.forces()
    ms["integrate"] += ctx.integrate()
print({k: round(v / a.steps, 4) for k, v in ms.items()}, "particles", sim.n, "overflow particles", ctx.counter("overflow_particles"))
