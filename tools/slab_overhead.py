"""How much does the slab machinery cost?  8,000,000 particles per GPU throughout (one rank's share of the 64M run):

  one process:   (a) the plain graph path, (b) a world-1 slab context (the slab loop: per-step host reads of the layer
                 starts, direct launches, boundary/interior split, but no neighbours), (c) the same, sequential variant
  under torchrun (N ranks, N GPUs): the tank is N shares long, every rank owns 8M particles and exchanges with its
                 neighbours over NCCL; device time, max over ranks

    python tools/slab_overhead.py [--steps 50] [--preroll 200]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/slab_overhead.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--preroll", type=int, default=200)
ap.add_argument("--share", type=float, nargs=3, default=[4.56, 4.56, 146.23 / 8.0])
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
out = {}
if world == 1:
    for mode in ("plain", "slab_world1", "slab_world1_sequential"):
        sim = gws.Simulator("cuda", tuple(a.share))
        if mode != "plain":
            sim.enable_slab(0, 1, bytes(128))
        sim.setup_scene()
        if mode == "slab_world1_sequential":
            sim.context().set_option("slab_overlap", 0)
        sim.step_many(a.preroll, timed=False)
        sim.step_many(5)
        ms = sim.step_many(a.steps)
        out[mode] = {"particles": sim.n, "ms_per_step": ms / a.steps}
        sim.close()
    print(json.dumps(out))
else:
    import torch
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ident = [gws.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    box = (a.share[0], a.share[1], a.share[2] * world)
    sim = gws.Simulator("cuda", box, device=local).enable_slab(rank, world, ident[0]).setup_scene()
    sim.step_many(a.preroll, timed=False)
    sim.step_many(5)
    dist.barrier()
    torch.cuda.synchronize()
    ms = sim.step_many(a.steps)
    t = torch.tensor([ms / a.steps], dtype=torch.float64, device="cuda")
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    if rank == 0:
        per_rank = [float(p.item()) for p in parts]
        print(json.dumps({f"slab_world{world}_nccl": {"particles_per_rank": sim.context().n, "ms_per_step_per_rank": per_rank,
                                                       "ms_per_step": max(per_rank)}}))
    dist.destroy_process_group()
