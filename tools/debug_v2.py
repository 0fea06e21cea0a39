import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gmu_water_simulation_b200 as gws
from oracle_binding import Oracle
box, steps = 0.4, 1
o = Oracle(box).setup_scene(); o.step(steps)
pos, vel = o.pos, o.vel
for variant in (0, 1):
    ctx = gws.SphContext(box, len(pos)); ctx.set_option("neighbour_variant", variant)
    ctx.upload(gws.particles_from_arrays(pos, vel))
    ctx.update_grid(); ctx.density_pressure()
    o2 = Oracle(box).set_state(pos, vel); o2.update_grid(); o2.update_density_pressure()
    rho, prs, _ = ctx.density_pressure_accel()
    gc, gl = ctx.neighbours(); oc, ol = o2.neighbours()
    rel = np.abs(rho / o2.density - 1)
    bad = np.argsort(-rel)[:8]
    print("variant", variant, "max rel", rel.max(), "count mismatches", int((gc != oc).sum()))
    keys = ctx.keys()
    for i in bad:
        k = keys[i]; cx, cy, cz = k % 9, (k // 9) % 9, k // 81
        f = (pos[i] + 0.2) / 0.0457 - np.array([cx, cy, cz])
        print(i, "rho gpu/ref", rho[i], o2.density[i], "cnt", gc[i], oc[i], "cell", (cx, cy, cz), "frac", f)
