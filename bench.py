#!/usr/bin/env python
"""bench.py — particle-steps/s of the SPH time step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full SPH step (grid build -> density/pressure -> forces -> walls+integration) over
the whole scene.  N=1 workload: dam break, box 3.62 -> 1,011,240 particles (BASELINE configs[2]).

* value  : whole-job particle-steps/s, state resident in HBM, timed with CUDA events on the stream
           the kernels are launched on (sph_step), L2 evicted before every timed step;
* e2e    : the same metric through the reference-facing plugin (CCUDAParticleSimulator::step) in
           RoundTrip mirror mode: every step uploads the 80-byte AoS host mirror, steps, and reads it
           back — the host<->device traffic of the reference's OpenCL path;
* roofline / cpu_baseline / clocks / gpu_launches as the driver contract asks.

--impl reference times the reference's own CPU implementation on the host: oracle/_ref/libsph_ref.so, the
reference's CCPUParticleSimulator sources compiled unmodified against stand-in Qt headers ("kind": "reference");
only where that file is missing (a checkout that never saw /root/reference) the bit-identical oracle port
("kind": "port").  Single-threaded, because the reference's CPU path is.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
ALG_BYTES_PER_PARTICLE_STEP = 260.0
KERNEL_ALG_BYTES = {"grid": 100.0, "density": 24.0, "forces": 56.0, "integrate": 80.0}  # per particle, §8d (+8 B cell arrays in grid)
WORKLOADS = {
    # name: (box, particles)
    "dam_break_1M": ((3.62, 3.62, 3.62), 1011240),
    "dam_break_250K": ((2.28, 2.28, 2.28), 250000),
    "dam_break_32K": ((1.14, 1.14, 1.14), 32500),
    "dam_break_16K": ((0.9, 0.9, 0.9), 16000),
    "tank_8M": ((4.56, 4.56, 146.23 / 8.0), 8000000),  # one GPU's share of the 64M tank, without slabs
}
# N > 1: BASELINE configs[4] — the elongated dam-break tank 4.56 x 4.56 x 146.23 with exactly 64,000,000
# particles (100 x 100 x 3200 cells), slab-decomposed along z over the N GPUs: the total work is fixed ("strong").
TANK_BOX = (4.56, 4.56, 146.23)
WORKLOADS["tank_64M"] = (TANK_BOX, 64000000)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# which hardware unit bounds each neighbour kernel (ncu) and the metric that shows it
BINDING_ROOF = {
    "k_density_mask": {"short": "issue+fp32", "name": "instruction issue / fp32 pipe (9.4 instructions per candidate test)",
                       "metric": "issue_active_frac"},
    "k_forces_mask": {"short": "l1", "name": "L1 data pipe: one wavefront per gathered neighbour record", "metric": "l1_data_pipe_frac"},
}


def ncu_facts(gws):
    """profiles/kernel_traffic.json, written by tools/kernel_traffic.py from an `ncu --set full` capture, is keyed by the
    hash of the kernel sources it was captured from; a capture of another build is refused (None)."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        facts = json.load(f)
    return facts if facts.get("kernel_source_sha256") == gws.kernel_source_hash() else None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(box, scenario=0):
    """The CPU implementation to time: the reference's own code (oracle/_ref) when built, else the oracle port
    (bit-identical to it, tests/test_ref_pins_oracle.py).  Returns (simulator, kind, description)."""
    import ref_binding

    cube = len(set(float(b) for b in box)) == 1
    if cube and ref_binding.available():
        return (ref_binding.Reference(box[0], scenario).setup_scene(), "reference",
                "oracle/_ref: the reference's CCPUParticleSimulator sources compiled unmodified (Qt stand-ins), 1 thread")
    from oracle_binding import Oracle, build_oracle

    build_oracle()
    why = "the reference's ctor only takes a cube" if not cube else "oracle/_ref is not built here"
    return (Oracle(box, scenario).setup_scene(), "port",
            f"oracle port of CCPUParticleSimulator, bit-identical to oracle/_ref on the test scenes ({why}), 1 thread")


def run_reference(args):
    """CPU arm: the reference's own CPU simulator on the host cores (single-threaded like the reference); rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # bounded sample of our arm's workload whose K+W steps fit in about two minutes of single-thread CPU time: N=1 the largest
    # dam-break cube that fits, N>1 a slice of the 64M tank (same cross-section and lattice, shorter in z)
    budget = 110.0 / max(args.steps + args.warmup, 1)
    per_particle_step = 4.0e-6
    if args.gpus > 1:
        workload, scaling = "tank_64M", "strong"
        frac = min(1.0, budget / per_particle_step / 64.0e6)
        cells_z = max(4, int(TANK_BOX[2] * frac / 0.0457))
        box = (TANK_BOX[0], TANK_BOX[1], cells_z * 0.0457)
        name = f"tank slice {box[0]} x {box[1]} x {box[2]:.4f}"
    else:
        workload, scaling = "dam_break_1M", "weak"
        name = "dam_break_16K"
        for cand in ("dam_break_1M", "dam_break_250K", "dam_break_32K", "dam_break_16K"):
            if WORKLOADS[cand][1] * per_particle_step <= budget:
                name = cand
                break
        if args.reference_sample:
            name = args.reference_sample
        box, _ = WORKLOADS[name]
    o, kind, what = cpu_arm(box)
    o.step(args.warmup)
    t0 = time.perf_counter()
    o.step(args.steps)
    sec = time.perf_counter() - t0
    value = o.n * args.steps / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": name, "particles": o.n, "note": what},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{name}: {o.n} particles x {args.steps} steps from the initial lattice (after {args.warmup} warm-up steps); "
                                   "the lattice has fewer neighbours per particle than the pre-rolled state our arm times, so this favours the CPU"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(pos, vel, box, seconds=12.0):
    """The CPU implementation timed on a bounded sample of the SAME state the GPU is benchmarked on."""
    from oracle_binding import Oracle

    o, kind, what = cpu_arm(box)
    if kind == "reference":
        o.set_state(pos, vel)  # same particle count as the scene (dam break): positions / velocities overwritten
    else:
        o = Oracle(box).set_state(pos, vel)
    o.step(1)  # the first step also pays for moving every particle to its cell
    n_steps = max(1, int(seconds / (o.n * 3.3e-6)))
    t0 = time.perf_counter()
    o.step(n_steps)
    sec = time.perf_counter() - t0
    out = {"value": o.n * n_steps / sec, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{o.n} particles (the benchmarked state) x {n_steps} step(s); {what}"}
    # extra, clearly NOT the reference's behaviour (its CPU path has no threading): the oracle port with its density and
    # force loops spread over all host cores (bit-identical results)
    o = Oracle(box).set_state(pos, vel)
    threads = o.set_threads(0)
    if threads > 1:
        o.step(1)
        t0 = time.perf_counter()
        o.step(n_steps)
        sec = time.perf_counter() - t0
        out["all_cores_variant"] = {"value": o.n * n_steps / sec, "unit": UNIT, "cores": threads,
                                    "note": "not reference behaviour: oracle port with density/force loops on all host cores"}
    return out


class JsonOut:
    """stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL, the CUDA
    runtime or a child process may print there), and only emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)

    def emit(self, line):
        self._out.write(json.dumps(line) + "\n")
        self._out.flush()


def bind_host_near_gpu(torch, local):
    """Pin this process to the CPU cores next to its GPU (NVML's affinity list) BEFORE any pinned host memory is
    allocated, so that the host mirror lands on that socket's memory (first touch).  What `numactl` / a job launcher
    does for a rank process; without it eight ranks moving 1.3 GB per step each share whatever socket they woke up
    on.  Returns what was done for the JSON line; never fatal."""
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = sorted(os.sched_getaffinity(0))
        return {"cpus_before": before, "cpus": len(after), "first_cpu": after[0], "last_cpu": after[-1], "how": "nvmlDeviceSetCpuAffinity"}
    except Exception as e:  # no NVML, restricted cpuset, ...: run unbound
        return {"unbound": f"{type(e).__name__}: {e}"[:120]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gmu_water_simulation_b200 as gws

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out = JsonOut()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    host_affinity = bind_host_near_gpu(torch, local) if args.bind_host else {"unbound": "--no-bind-host"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line, whatever NCCL_DEBUG says
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    slab = world > 1
    if slab:
        workload = "tank_64M"
        box = TANK_BOX
        ident = [gws.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sim = gws.Simulator("cuda", box, device=local).enable_slab(rank, world, ident[0]).setup_scene()
        sim.context().set_option("slab_rebalance", args.slab_rebalance)
    else:
        workload = args.workload
        box, _ = WORKLOADS[workload]
        sim = gws.Simulator("cuda", box, device=local).setup_scene()
    ctx = sim.context()
    n = sim.n
    if args.neighbour_variant is not None:
        ctx.set_option("neighbour_variant", args.neighbour_variant)

    # ---- pre-roll so the dam has collapsed and the state is irregular, then warm-up
    sim.step_many(1, timed=False)
    sim.step_many(max(args.preroll - 1, 1), timed=False)
    ctx.synchronize()
    if not slab:
        ctx.state_save(1)  # slot 1: the state after the pre-roll (start of the fixed window run further down)
    flush_l2 = args.flush_l2 and not slab  # a slab's working set (GBs) is far larger than L2 anyway
    args.flush_l2 = flush_l2
    ctx.set_option("flush_l2", 1 if flush_l2 else 0)
    warmup = max(args.warmup, 3)
    sim.step_many(warmup)

    def mean_neighbours():
        return ctx.counter("neighbour_pairs") / max(ctx.n, 1)  # of the last density pass, self included

    nb_start = None
    if not slab:
        ctx.state_save(0)  # slot 0: the state at the start of the timed window (replayed for the per-kernel split)
        nb_start = mean_neighbours()

    # ---- timed region: exactly K steps, device time from CUDA events on the launching stream
    launches0 = ctx.counter("kernel_launches")
    barrier()
    with ClockSampler(local) as clocks:
        wall0 = time.perf_counter()
        dev_ms = sim.step_many(args.steps)
        barrier()
        wall = time.perf_counter() - wall0
        launches = ctx.counter("kernel_launches") - launches0
        nb_end = None if slab else mean_neighbours()
        if wall < 1.5 and not slab:  # keep the GPU under the same load long enough for a few clock samples
            ctx.set_option("flush_l2", 0)
            t_end = time.perf_counter() + 1.5
            while time.perf_counter() < t_end:
                sim.step_many(50, timed=True)
            ctx.set_option("flush_l2", 1 if args.flush_l2 else 0)
    dev_ms = max_over_ranks(dev_ms)
    total_particles = sum_over_ranks(float(ctx.n if slab else n))
    value = total_particles * args.steps / (dev_ms * 1e-3)
    if slab:
        finish_slab(args, sim, ctx, dist, rank, world, local, box, workload, value, dev_ms, wall, launches, total_particles,
                    clocks, barrier, max_over_ranks, gws, out, host_affinity)
        return

    # ---- per-kernel split ON THE SAME STEPS: the window is replayed from the saved state (same particle order in
    # memory, deterministic arithmetic => identical steps) with CUDA events between the kernel groups
    ctx.state_restore(0)
    prof = ctx.step_profiled(args.steps)
    phase_ms = {k: prof[k] / args.steps for k in ("grid", "density", "forces")}
    phase_ms["sum"] = sum(phase_ms.values())
    phase_ms["step_by_events"] = prof["total"] / args.steps

    # ---- the same window with a warm L2 (no eviction), for information
    ctx.set_option("flush_l2", 0)
    ctx.state_restore(0)
    warm_ms = max_over_ranks(sim.step_many(args.steps))
    ctx.set_option("flush_l2", 1 if args.flush_l2 else 0)

    # ---- fixed window, independent of --steps / --warmup: steps preroll+10 .. preroll+210 of the same run
    ctx.state_restore(1)
    sim.step_many(10)
    fixed_steps = 200
    fixed_ms = sim.step_many(fixed_steps)
    nb_fixed_end = mean_neighbours()
    fixed_window = {"value": n * fixed_steps / (fixed_ms * 1e-3), "unit": UNIT, "ms_per_step": fixed_ms / fixed_steps,
                    "steps": [args.preroll + 10, args.preroll + 10 + fixed_steps], "mean_neighbours_at_end": nb_fixed_end,
                    "note": "same run, L2 evicted before every step; the cost of a step grows as the water piles up, so this "
                            "window does not move with --steps / --warmup"}
    ctx.state_restore(0)  # everything below (e2e, CPU baseline) starts from the state the timed window started from

    peak, peak_src = measured_peaks()
    top = max(("density", "forces"), key=lambda k: phase_ms[k])
    kernel_name = {"density": "k_density_mask", "forces": "k_forces_mask"}[top]
    ncu = ncu_facts(gws)  # dram bytes + pipe utilisations per kernel from the committed ncu capture of THIS build, or None
    facts = (ncu or {}).get("kernels", {}).get(kernel_name) if workload == "dam_break_1M" else None
    top_bytes = KERNEL_ALG_BYTES[top] * n
    top_gbs = top_bytes / (phase_ms[top] * 1e-3) / 1e9
    step_gbs = value / world * ALG_BYTES_PER_PARTICLE_STEP / 1e9
    binding = BINDING_ROOF[kernel_name]
    roofline = {
        # the contract's HBM figure: ALGORITHMIC bytes / in-window kernel time against the measured copy bandwidth ...
        "bound": binding["short"], "kernel": kernel_name, "achieved": top_gbs, "peak": peak, "unit": "GB/s", "frac": top_gbs / peak,
        "traffic": facts["dram_bytes"] if facts else None, "peak_source": peak_src,
        "alg_bytes_per_launch": top_bytes, "kernel_ms": phase_ms[top],
        # ... and the roof that actually binds this kernel (ncu), because HBM does not
        "binding_roof": {"name": binding["name"], "frac": (facts or {}).get(binding["metric"]), "unit": "fraction of peak (ncu)",
                         "source": (ncu or {}).get("source") if facts else "no ncu capture of this build committed (profiles/kernel_traffic.json is keyed by the hash of the kernel sources)"},
        "step_level": {"alg_bytes_per_particle_step": ALG_BYTES_PER_PARTICLE_STEP, "achieved": step_gbs, "frac": step_gbs / peak},
        "note": "neither neighbour kernel is HBM-bound: achieved/peak/frac above are the HBM roofline the north star asks for, "
                "binding_roof is the unit ncu shows saturated (DESIGN.md section 3.1)",
    }
    mean_nb = 0.5 * (nb_start + nb_end)

    # ---- end to end through the plugin: upload AoS mirror + step + download, every step
    e2e = None
    if rank == 0 or world > 1:
        sim.set_mirror_mode(2)
        sim.step(2)
        barrier()
        t0 = time.perf_counter()
        sim.step(args.e2e_steps)
        barrier()
        e2e_sec = max_over_ranks(time.perf_counter() - t0)
        sim.set_mirror_mode(0)
        e2e = {"value": total_particles * args.e2e_steps / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": 80 * n * world,
               "d2h_bytes_per_step": 80 * n * world, "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_sec / args.e2e_steps,
               "path": "CCUDAParticleSimulator::step(), MirrorMode::RoundTrip (pinned 80-byte AoS host mirror up and down every step)"}
        # for information: the viewer bridge (state resident, the 80-byte records of every stride-th step delivered
        # to the host mirror by an overlapped copy); not the headline, there is no host->device traffic on this path
        viewer = {}
        for stride in (1, 4):
            sim.set_mirror_mode(3)
            sim.set_mirror_stride(stride)
            sim.step(2 * stride)
            sim.wait_host()
            barrier()
            t0 = time.perf_counter()
            sim.step(args.e2e_steps)
            sim.wait_host()
            barrier()
            sec = max_over_ranks(time.perf_counter() - t0)
            viewer[f"stride{stride}"] = total_particles * args.e2e_steps / sec
        sim.set_mirror_mode(0)
        sim.set_mirror_stride(1)
        e2e["viewer_async_download"] = dict(viewer, unit=UNIT, d2h_bytes_per_refresh=80 * n * world,
                                            path="MirrorMode::AsyncDownload (sph_download_particles_async)")

    device_name = sim.device
    grid_res = list(ctx.grid_res)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sim.sync_host()
        hp = sim.host_particles()
        cpu = cpu_baseline_sample(hp["position"][:, :3].copy(), hp["velocity"][:, :3].copy(), box, args.cpu_seconds)

    scaling_baseline = None
    if args.scaling_baseline and workload != "tank_64M":
        # the N>1 runs decompose the fixed 64M tank; this is the same tank on this one GPU (no slabs)
        del sim, ctx
        big = gws.Simulator("cuda", TANK_BOX, device=local).setup_scene()
        big.step_many(1, timed=False)
        big.step_many(max(args.preroll - 1, 1), timed=False)
        big.step_many(max(args.warmup, 3))
        big_steps = args.steps  # same pre-roll, warm-up and timed window as the N>1 runs: the step cost grows as the dam collapses
        big_ms = big.step_many(big_steps)
        scaling_baseline = {"workload": "tank_64M", "particles": big.n, "value": big.n * big_steps / (big_ms * 1e-3),
                            "unit": UNIT, "ms_per_step": big_ms / big_steps, "steps": big_steps,
                            "note": "denominator for the strong-scaling series run with --gpus 2/4/8"}
        device_name = big.device
        del big
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "particles_per_gpu": n, "box": list(box), "grid": grid_res,
                       "preroll_steps": args.preroll, "timed_window_steps": [args.preroll + warmup, args.preroll + warmup + args.steps],
                       "mean_neighbours": mean_nb, "mean_neighbours_window": [nb_start, nb_end],
                       "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (slab mode pending)",
                       "l2": "L2 evicted (256 MiB scratch write, then read back so that no dirty scratch lines remain) before every timed step" if args.flush_l2 else "no eviction"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "host_affinity": host_affinity,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "phase_ms": phase_ms,
            "fixed_window": fixed_window,
            "value_warm_l2": total_particles * args.steps / (warm_ms * 1e-3),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "device": device_name,
            "scaling_baseline": scaling_baseline,
        }
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def state_checksum(rec):
    """Order-independent 64-bit checksum of (id, position bits, velocity bits) over a set of particle records."""
    import numpy as np

    h = rec["id"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for field, mult in (("position", 0xC2B2AE3D27D4EB4F), ("velocity", 0x165667B19E3779F9)):
        bits = np.ascontiguousarray(rec[field][:, :3]).view(np.uint32).astype(np.uint64)
        for axis in range(3):
            h ^= (bits[:, axis] + np.uint64(axis + 1)) * np.uint64(mult)
            h = (h << np.uint64(13)) | (h >> np.uint64(51))
    return int(h.sum(dtype=np.uint64))


def slab_equivalence(gws, dist, rank, world, local, steps=10):
    """SURVEY.md section 8e on the bench box itself: the 4,000,000-particle tank 4.56 x 4.56 x 9.13 stepped by `world`
    slabs over NCCL and by one GPU; the merged slab state must be bit-identical (checksum over ids, positions,
    velocities of all particles)."""
    box = (4.56, 4.56, 9.13)
    ident = [gws.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    sim = gws.Simulator("cuda", box, device=local).enable_slab(rank, world, ident[0]).setup_scene()
    sim.step_many(steps)
    owned = sim.context().download_owned()
    mine = (int(owned.shape[0]), state_checksum(owned), int(sim.context().counter("slab_far_movers")))
    sim.close()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    out = None
    if rank == 0:
        plain = gws.Simulator("cuda", box, device=local).setup_scene()
        plain.step_many(steps)
        rec = plain.context().download_owned()
        want = state_checksum(rec)
        got = sum(p[1] for p in parts) & 0xFFFFFFFFFFFFFFFF
        out = {"tank": "4.56 x 4.56 x 9.13", "particles": int(rec.shape[0]), "steps": steps, "slabs": world,
               "particles_in_slabs": sum(p[0] for p in parts), "far_movers": sum(p[2] for p in parts),
               "checksum_one_gpu": f"{want:016x}", "checksum_slabs": f"{got:016x}", "bitwise_identical": got == want}
        plain.close()
    return out


def finish_slab(args, sim, ctx, dist, rank, world, local, box, workload, value, dev_ms, wall, launches, total_particles,
                clocks, barrier, max_over_ranks, gws, out, host_affinity):
    """N > 1: e2e (upload + step + read-back of the owned particles every step), per-rank slab facts, the JSON line."""
    peak, peak_src = measured_peaks()
    sim.set_mirror_mode(2)  # RoundTrip: every rank uploads its owned 80-byte records, steps, reads them back
    sim.step(2)
    barrier()
    t0 = time.perf_counter()
    sim.step(args.e2e_steps)
    barrier()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    sim.set_mirror_mode(0)
    info = ctx.slab_info()
    info["face_moves"] = ctx.counter("slab_face_moves")
    info["load_us"] = ctx.counter("slab_load_us")
    info["far_movers"] = ctx.counter("slab_far_movers")  # particles the boundary-only exchange would have missed: must be 0
    info["clocks"] = clocks.summary()  # every rank samples its own GPU: the step runs at the pace of the slowest slab
    info["host_affinity"] = host_affinity
    infos = [None] * world
    dist.all_gather_object(infos, info)
    sim.close()
    equivalence = None if args.no_equivalence else slab_equivalence(gws, dist, rank, world, local)
    if rank == 0:
        step_gbs = value / world * ALG_BYTES_PER_PARTICLE_STEP / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "particles": int(total_particles), "box": list(box),
                       "preroll_steps": args.preroll,
                       "parallelism": f"{world} z-slabs, ghost+migration exchange per step via ncclSend/ncclRecv; "
                                      + ("faces follow the measured load" if args.slab_rebalance else "static equal-layer faces"),
                       "slabs": [{k: i[k] for k in ("z0", "z1", "n_own", "mean_sent_per_step", "far_movers", "face_moves", "load_us")} for i in infos],
                       "particles_conserved": int(sum(i["n_own"] for i in infos)) == WORKLOADS["tank_64M"][1],
                       "l2": "no eviction: the per-GPU working set (GBs) is far larger than the 126 MB L2"},
            "clocks": clocks.summary(),
            "clocks_per_rank": [i["clocks"] for i in infos],
            "host_affinity_per_rank": [i["host_affinity"] for i in infos],
            "e2e": {"value": total_particles * args.e2e_steps / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(80 * total_particles),
                    "d2h_bytes_per_step": int(80 * total_particles), "steps": args.e2e_steps,
                    "ms_per_step": 1e3 * e2e_sec / args.e2e_steps,
                    "path": "CCUDAParticleSimulator::step() per rank, MirrorMode::RoundTrip (owned 80-byte records uploaded before and read back after every step)"},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "issue+fp32 / l1 (per-kernel split at N=1)", "kernel": "whole step (per GPU)", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                         "frac": step_gbs / peak, "traffic": None, "peak_source": peak_src,
                         "note": "step-level: 260 algorithmic B per particle-step (SURVEY.md \u00a78d); per-kernel split is reported at N=1"},
            "cpu_baseline": None,
            "slab_equivalence": equivalence,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "device": infos and sim_device_name(gws, local),
        }
        out.emit(line)
    dist.destroy_process_group()


def sim_device_name(gws, local):
    try:
        return gws.device_name(local)
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dam_break_1M", choices=sorted(WORKLOADS))
    ap.add_argument("--reference-sample", default=None, choices=sorted(WORKLOADS),
                    help="--impl reference: force the sampled scene instead of sizing it to the time budget")
    ap.add_argument("--preroll", type=int, default=200)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-baseline", dest="scaling_baseline", action="store_false",
                    help="skip the single-GPU run of the 64M tank (denominator of the strong-scaling series)")
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false")
    ap.add_argument("--slab-rebalance", type=int, default=1, choices=[0, 1],
                    help="N>1: 1 = the faces between slabs follow the measured load (default), 0 = static equal-layer split")
    ap.add_argument("--no-equivalence", action="store_true", help="N>1: skip the 4M-tank slabs == one GPU checksum run")
    ap.add_argument("--no-bind-host", dest="bind_host", action="store_false",
                    help="do not pin the process to the CPU cores next to its GPU")
    ap.add_argument("--neighbour-variant", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
