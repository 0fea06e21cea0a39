import os, sys
sys.path.insert(0, "/root/repo")
import gmu_water_simulation_b200 as gws
sim = gws.Simulator("cuda", 3.62).setup_scene(); ctx = sim.context()
ctx.set_option("use_graph", 0)
ctx.step(50, timed=False); ctx.synchronize()
for k in range(4):
    print([round(x, 4) for x in (ctx.update_grid(), ctx.density_pressure(), ctx.forces(), ctx.collisions(), ctx.integrate())])
