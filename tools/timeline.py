"""In-situ kernel timeline of a few steps (CUPTI activity records through torch.profiler): real durations inside the
CUDA graph and the gaps between the kernels, which an ncu launch list (serialised, cold cache) cannot show."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
import gmu_water_simulation_b200 as gws

box = float(sys.argv[1]) if len(sys.argv) > 1 else 3.62
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
use_graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.zeros(1, device="cuda")
sim = gws.Simulator("cuda", box).setup_scene()
ctx = sim.context()
ctx.set_option("use_graph", use_graph)
ctx.step(1, timed=False); ctx.step(199, timed=False); ctx.synchronize()
ctx.step(5)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ms = ctx.step(steps)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
print(f"{steps} steps, {ms / steps * 1e3:.1f} us/step by CUDA events; {len(ev)} device activities")
t_prev_end = None
busy = 0.0
for e in ev:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = (s - t_prev_end) if t_prev_end is not None else 0.0
    print(f"{e.name[:40]:40s} start {s - ev[0].time_range.start:10.1f} us  dur {d:8.1f} us  gap {gap:6.1f} us")
    t_prev_end = e.time_range.end
    busy += d
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us ({100 * (span - busy) / span:.1f} %)")
