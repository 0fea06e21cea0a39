"""Why did the CPU oracle sample get slower after the async-mirror section of bench.py?  Time one oracle step on the
device state before and after stepping in MirrorMode 3 / 1, and print the state statistics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gmu_water_simulation_b200 as gws
from oracle_binding import Oracle

box = 3.62
sim = gws.Simulator("cuda", box).setup_scene()
ctx = sim.context()
sim.step_many(1, timed=False); sim.step_many(249, timed=False)

def probe(tag):
    sim.sync_host()
    hp = sim.host_particles()
    pos, vel = hp["position"][:, :3].copy(), hp["velocity"][:, :3].copy()
    ids_ok = bool(np.array_equal(hp["id"], np.arange(len(hp), dtype=np.uint32)))
    o = Oracle(box).set_state(pos, vel)
    o.step(1)
    t0 = time.perf_counter(); o.step(1); dt = time.perf_counter() - t0
    counts, _ = o.neighbours(lists=False) if hasattr(o, "neighbours") else (None, None)
    print(tag, f"oracle {dt:.2f} s/step, ids_by_index={ids_ok}, finite={bool(np.isfinite(pos).all())}, "
          f"|v|max={np.linalg.norm(vel, axis=1).max():.2f}, stats={ctx.stats()}", flush=True)

probe("A after 250 fused steps      ")
sim.set_mirror_mode(1); sim.step(10); sim.set_mirror_mode(0)
probe("B after 10 Download steps    ")
sim.set_mirror_mode(3); sim.set_mirror_stride(1); sim.step(10); sim.wait_host(); sim.set_mirror_mode(0)
probe("C after 10 AsyncDownload steps")
