"""In-situ timeline of a few slab-mode steps on rank 0 (CUPTI records through torch.profiler), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/timeline_slab.py [--box 4.56 4.56 36.56] [--steps 3]

Shows what an ncu launch list cannot: which kernels of the overlapped step run concurrently with the exchange and
where the device idles."""
import argparse, os, sys
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--box", type=float, nargs=3, default=[4.56, 4.56, 36.56])
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--preroll", type=int, default=200)
ap.add_argument("--no-overlap", action="store_true")
ap.add_argument("--print-rank", type=int, default=0, help="whose timeline to print")
ap.add_argument("--summary", action="store_true", help="every rank: per-kernel time per step, span and idle time, SM clock while stepping")
a = ap.parse_args()
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
torch.zeros(1, device="cuda")
ident = [gws.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ident, src=0)
sim = gws.Simulator("cuda", tuple(a.box), device=local).enable_slab(rank, world, ident[0]).setup_scene()
ctx = sim.context()
if a.no_overlap:
    ctx.set_option("slab_overlap", 0)
sim.step_many(a.preroll, timed=False)
ms = sim.step_many(10)
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t = sim.step_many(a.steps)
    torch.cuda.synchronize()
dist.barrier()
if a.summary:
    import collections, subprocess
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    per = collections.defaultdict(float)
    cover_end, idle = ev[0].time_range.start, 0.0
    for e in ev:
        name = e.name.split("(")[0].replace("void ", "").replace("sph::", "")
        per[name] += (e.time_range.end - e.time_range.start) / a.steps
        idle += max(e.time_range.start - cover_end, 0.0)
        cover_end = max(cover_end, e.time_range.end)
    # clocks while the same loop keeps running
    import threading
    samples = []
    stop = threading.Event()
    def smi():
        while not stop.is_set():
            out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_slowdown", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
            if out:
                samples.append(out)
            stop.wait(0.1)
    th = threading.Thread(target=smi); th.start()
    t_long = sim.step_many(300)
    stop.set(); th.join()
    mine = {"rank": rank, "particles": ctx.n, "us_per_step_before": ms / 10 * 1e3, "us_per_step_profiled": t / a.steps * 1e3,
            "us_per_step_300": t_long / 300 * 1e3, "span_us_per_step": (cover_end - ev[0].time_range.start) / a.steps,
            "idle_us_per_step": idle / a.steps, "kernels_us_per_step": {k: round(v, 1) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
            "smi": samples[:: max(len(samples) // 6, 1)]}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        import json
        for r in allr:
            print(json.dumps(r))
    dist.barrier()
    sys.exit(0)
if rank == a.print_rank:
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    print(f"rank {rank} of {world}: {ctx.n} local particles, {ms / 10 * 1e3:.1f} us/step before profiling, {t / a.steps * 1e3:.1f} us/step profiled")
    t0 = ev[0].time_range.start
    cover_end = t0
    idle = 0.0
    for e in ev:
        s, en = e.time_range.start, e.time_range.end
        gap = s - cover_end
        if gap > 0:
            idle += gap
        print(f"{e.name[:44]:44s} start {s - t0:9.1f} us  dur {en - s:8.1f} us  {'idle before %.1f us' % gap if gap > 1.0 else ''}")
        cover_end = max(cover_end, en)
    span = cover_end - t0
    print(f"span {span:.1f} us over {a.steps} steps, device idle {idle:.1f} us ({100 * idle / span:.1f} %)")
dist.barrier()
