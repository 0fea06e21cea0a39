"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gmu_water_simulation_b200 as gws

box = 0.6
sim = gws.Simulator("cuda", box).setup_scene()
ctx = sim.context()
sim.step_many(6)                      # fused graph path
ctx.set_option("use_graph", 0)
sim.step_many(2)                      # direct launches
ctx.update_grid(); ctx.density_pressure(); ctx.forces(); ctx.collisions(); ctx.integrate()   # phase path
ctx.update_grid(); ctx.density_pressure()
ctx.keys(); ctx.permutation(); ctx.cell_start(); ctx.neighbours(); ctx.density_pressure_accel(); ctx.stats()
ctx.forces(); ctx.integrate(); ctx.download()
ctx.set_option("neighbour_variant", 0)
ctx.step(2)
ctx.set_option("neighbour_variant", 1)
ctx.step(2)
ctx.update_grid(); ctx.density_pressure()
ctx.neighbours(source="mask")                              # k_mask_lists: the stored hit words decoded
ctx.fill_height_percentile(0.95)                           # k_minmax_y, k_hist_y
ctx.counter("neighbour_pairs")                             # k_sum_i32
ctx.forces(); ctx.integrate()
ctx.state_save(0); ctx.step(2); ctx.state_restore(0); ctx.step_profiled(2)
ctx.brute_density_pressure(); ctx.brute_forces(); ctx.integrate()
# dense clump -> overflow path
rng = np.random.default_rng(1)
pos = (rng.random((1500, 3), dtype=np.float32) - 0.5) * np.float32(0.03)
c2 = gws.SphContext(0.4, 1500); c2.upload(gws.particles_from_arrays(pos)); c2.step(2)
c2.upload(gws.particles_from_arrays(pos))                  # the clump again (two steps blew it apart)
c2.update_grid(); c2.density_pressure(); c2.neighbours(source="mask"); c2.forces(); c2.integrate()   # overflow list decoded too
print("overflow particles", c2.counter("overflow_particles"))
# collision mesh (planes looped in the fused epilogue and in k_integrate_collide), second density pass on one grid
b = np.float32(0.2)
face = [-0.5, -0.5, 0.0, -b + 0.04, -b, -b, -b, -b + 0.04, -b, -b + 0.04, -b, b]
m = gws.Simulator("cuda", 0.4).setup_scene(); m.set_collision_faces(np.array([face], dtype=np.float32)); m.step_many(4)
mc = m.context(); mc.update_grid(); mc.density_pressure(); mc.density_pressure(); mc.forces(); mc.integrate()
# fountain (append) and single-rank slab mode (pack kernel, range launches)
f = gws.Simulator("cuda", 0.4, scenario=gws.FOUNTAIN).setup_scene(); f.step(20)
f.step_many(30)                                            # device-side emitter (k_emit) + fused step
s = gws.Simulator("cuda", (0.4, 0.4, 0.9)).enable_slab(0, 1, bytes(128)).setup_scene(); s.step_many(5)
# two slab ranks in this process over the loop-back transport (pack, exchange by device-to-device copies, overlap)
import threading
ident = gws.comm_local_id(2)
ranks = [gws.Simulator("cuda", (0.4, 0.4, 0.9)).enable_slab(r, 2, ident).setup_scene() for r in range(2)]
th = [threading.Thread(target=lambda q=q: q.step_many(4)) for q in ranks]
[t.start() for t in th]; [t.join() for t in th]
print("sanitize run ok", sim.n, f.n, s.context().n, [q.context().n for q in ranks])
