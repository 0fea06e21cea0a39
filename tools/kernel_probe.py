"""Per-kernel A/B timing on the benchmarked state (1,011,240 particles, 200 steps into the dam break).

    python tools/kernel_probe.py [--tunings 0 2 ...] [--reps 20]

For each value of the "tuning" option the phases are timed with CUDA events through the C ABI, L2 evicted before every
repetition (same method as bench.py's phase_ms).  Development tool: prints one JSON line per tuning value.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import gmu_water_simulation_b200 as gws

    ap = argparse.ArgumentParser()
    ap.add_argument("--tunings", type=int, nargs="+", default=[0])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--box", type=float, default=3.62)
    ap.add_argument("--preroll", type=int, default=200)
    a = ap.parse_args()
    sim = gws.Simulator("cuda", a.box).setup_scene()
    ctx = sim.context()
    sim.step_many(a.preroll)
    sim.sync_host()
    hp = sim.host_particles()
    rec = hp.copy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for tuning in a.tunings:
        ctx.set_option("tuning", tuning)
        ctx.upload(rec)                      # every variant starts from the same state
        ms = {"grid": 0.0, "density": 0.0, "forces": 0.0, "integrate": 0.0}
        for rep in range(a.reps + 2):
            ctx.upload(rec)
            flush.zero_()
            flush.view(torch.int32).sum()
            torch.cuda.synchronize()
            t = (ctx.update_grid(), ctx.density_pressure(), ctx.forces(), ctx.integrate())
            if rep >= 2:
                for k, v in zip(ms, t):
                    ms[k] += v / a.reps
        ctx.upload(rec)
        fused = ctx.step(a.reps) / a.reps     # graph path, warm L2
        print(json.dumps({"tuning": tuning, **{k: round(v, 4) for k, v in ms.items()}, "fused_step_warm_ms": round(fused, 4)}), flush=True)


if __name__ == "__main__":
    main()
