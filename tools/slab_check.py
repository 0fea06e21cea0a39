"""Slab-mode equivalence check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_check.py [--box 0.5 0.5 1.5] [--steps 30]

Every rank steps its z-slab (ghost exchange + migration over NCCL); rank 0 merges the owned particles by
id and compares with a single-GPU run of the same tank: ids form a partition, cell ids exact after the
first step, positions within the fp32 tolerance (summation order inside the density pass depends on the
4-alignment of the local array, so low bits may differ between decompositions).
"""
import argparse
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmu_water_simulation_b200 as gws  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--box", type=float, nargs=3, default=[0.5, 0.5, 1.5])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--no-overlap", action="store_true", help="sequential exchange-then-step instead of the overlapped step")
    ap.add_argument("--bitwise", action="store_true", help="require positions, velocities, densities bit-identical to the single-GPU run")
    ap.add_argument("--gz", type=float, default=0.0, help="gravity z component (pushes water across the slab faces)")
    a = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    def log(msg):
        print(f"[rank {os.environ.get('RANK', '?')}] {msg}", file=sys.stderr, flush=True)

    log("init gloo")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    ident = [gws.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    log("nccl id broadcast done; creating slab simulator")

    sim = gws.Simulator("cuda", tuple(a.box), device=local).enable_slab(rank, world, ident[0]).setup_scene()
    ctx = sim.context()
    if a.gz:
        sim.set_gravity((0.0, -9.80665, a.gz))
    if a.no_overlap:
        ctx.set_option("slab_overlap", 0)
    info0 = ctx.slab_info()
    log(f"slab ready {info0}")
    checks = {}

    def gather_state():
        rec = ctx.download_owned()
        parts = [None] * world if rank == 0 else None
        dist.gather_object(rec, parts, dst=0)
        if rank != 0:
            return None
        allrec = np.concatenate(parts)
        order = np.argsort(allrec["id"], kind="stable")
        return allrec[order]

    ref = None
    if rank == 0:
        ref = gws.Simulator("cuda", tuple(a.box), device=local).setup_scene()
        if a.gz:
            ref.set_gravity((0.0, -9.80665, a.gz))
    for k, nsteps in enumerate([1, a.steps - 1]):
        sim.step_many(nsteps)
        log(f"stepped {nsteps}")
        merged = gather_state()
        if rank == 0:
            ref.step_many(nsteps)
            ref.sync_host()
            hp = ref.host_particles()
            n = ref.n
            assert merged.shape[0] == n, f"slab ranks hold {merged.shape[0]} particles, the tank has {n}"
            assert np.array_equal(merged["id"], np.arange(n, dtype=np.uint32)), "owned sets are not a partition of the ids"
            if a.bitwise:
                for f in ("cell_id", "position", "velocity", "density", "pressure", "acceleration"):
                    same = merged[f].view(np.uint32) == hp[f].view(np.uint32)
                    assert same.all(), f"{f}: {(~same).sum()} words differ from the single-GPU run after step {1 if k == 0 else a.steps}"
                checks[f"bitwise_after_{1 if k == 0 else a.steps}_steps"] = True
            err = np.abs(merged["position"][:, :3] - hp["position"][:, :3]).max()
            checks[f"max_abs_dx_after_{1 if k == 0 else a.steps}_steps"] = float(err)
            if k == 0:
                assert np.array_equal(merged["cell_id"], hp["cell_id"]), "cell ids differ from the single-GPU run"
                assert np.array_equal(merged["density"] > 0, hp["density"] > 0)
                rel = np.abs(merged["density"] / hp["density"] - 1).max()
                checks["max_rel_density_step1"] = float(rel)
                assert rel <= 1e-5 and err <= 1e-6, (rel, err)
            else:
                # SPH is chaotic: after many steps only statistics are comparable
                d = np.abs(merged["position"][:, :3] - hp["position"][:, :3]).max(axis=1)
                checks["p95_abs_dx"] = float(np.percentile(d, 95))
                com = np.abs(merged["position"][:, :3].mean(axis=0) - hp["position"][:, :3].mean(axis=0)).max()
                checks["com_abs_diff"] = float(com)
                assert np.percentile(d, 95) <= 0.1 * 0.0457 and err <= 0.0457 and com <= 1e-4, checks
    far = ctx.counter("slab_far_movers")
    assert far == 0, f"{far} particles crossed more than 2 z-layers in one step"
    infos = [None] * world if rank == 0 else None
    dist.gather_object(ctx.slab_info(), infos, dst=0)
    if rank == 0:
        print("SLAB CHECK OK", {"world": world, "particles": int(ref.n), "checks": checks, "slabs": infos, "first": info0})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback

        traceback.print_exc()
        os._exit(1)  # never leave the other ranks waiting in a collective
