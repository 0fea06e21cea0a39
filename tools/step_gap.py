"""Where does the step time go beyond the sum of the kernel times?  Warm-L2 timings of the same 100 steps as
(a) CUDA-graph steps, (b) direct launches, (c) phase calls without intermediate syncs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws

def fresh():
    sim = gws.Simulator("cuda", 3.62).setup_scene()
    ctx = sim.context()
    ctx.step(1, timed=False); ctx.step(199, timed=False); ctx.synchronize()
    return sim, ctx

sim, ctx = fresh()
ctx.step(3)
print("graph      ms/step", ctx.step(100) / 100)
sim, ctx = fresh()
ctx.set_option("use_graph", 0)
ctx.step(3)
print("direct     ms/step", ctx.step(100) / 100)
sim, ctx = fresh()
ctx.set_option("use_graph", 0); ctx.set_option("fuse_integrate", 0)
ctx.step(3)
print("direct, unfused ms/step", ctx.step(100) / 100)
sim, ctx = fresh()
ph = {"grid": 0.0, "density": 0.0, "forces": 0.0, "integrate": 0.0}
for _ in range(100):
    ph["grid"] += ctx.update_grid(); ph["density"] += ctx.density_pressure(); ph["forces"] += ctx.forces(); ph["integrate"] += ctx.integrate()
print("phases (each timed, synced) ms", {k: round(v / 100, 4) for k, v in ph.items()}, "sum", round(sum(ph.values()) / 100, 4))
