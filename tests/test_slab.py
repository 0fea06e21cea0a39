"""Slab decomposition (multi-GPU extension): host logic on CPU (incl. a world_size-2 gloo run) and the
device path on however many GPUs the box has."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_binding import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("rz,world", [(33, 2), (3200, 8), (800, 2), (17, 4), (101, 3)])
def test_slab_plan_is_a_partition(gws, rz, world):
    edges = [gws.slab_plan(rz, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == rz
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a1 - a0 >= 4
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_slab_plan_rejects_thin_slabs(gws):
    with pytest.raises(gws.SphError):
        gws.slab_plan(15, 4, 0)
    with pytest.raises(gws.SphError):
        gws.slab_plan(100, 4, 4)


def test_face_shift_rule(gws):
    """The load-balance rule is a pure function that both ranks at a face evaluate on the same two records."""
    f = gws.slab_face_shift
    lo, hi = [4000, 0, 400, 1], [3600, 400, 800, 1]          # {load_us, z0, z1, may_grow}
    assert f(lo, hi) == -1                                    # the lower rank is busier: it gives its top layer away
    assert f(hi[:1] + lo[1:], lo[:1] + hi[1:]) == +1          # loads swapped: the face moves up
    assert f([4000, 0, 400, 1], [3990, 400, 800, 1]) == 0     # within the 0.4 % hysteresis
    assert f(lo, [3600, 400, 800, 0]) == 0                    # the receiver may not grow
    assert f(lo, hi, shift=-8, shift_max=8) == 0              # already shift_max layers below the plan
    assert f([4000, 0, 10, 1], hi) == 0                       # the giving slab is at the minimum thickness
    assert f([0, 0, 400, 1], hi) == 0 and f(lo, [0, 400, 800, 1]) == 0   # no load measured yet
    assert f(lo, hi, mode=0) == 0
    # test pattern: three exchanges up, three down, phase-shifted by the face index; limits still apply
    assert [f(lo, hi, mode=2, exchange=e, face=0) for e in range(6)] == [1, 1, 1, -1, -1, -1]
    assert [f(lo, hi, mode=2, exchange=e, face=1) for e in range(6)] == [1, 1, -1, -1, -1, 1]
    assert f(lo, hi, shift=8, shift_max=8, mode=2, exchange=0) == 0


def test_scene_partition_matches_oracle(gws):
    """Each rank keeps exactly the particles whose z-layer (the oracle's key formula) it owns; ids stay global."""
    box = (0.5, 0.3, 1.2)
    o = Oracle(box).setup_scene()
    rz = o.grid_res[2]
    layer = o.keys() // (o.grid_res[0] * o.grid_res[1])  # all particles are still in their lattice positions
    seen = np.zeros(o.n, dtype=np.int32)
    for rank in range(3):
        z0, z1 = gws.slab_plan(rz, 3, rank)
        sim = gws.Simulator("scene_only", box).set_owned_layers(z0, z1).setup_scene()
        hp = sim.host_particles()
        ids = hp["id"].astype(np.int64)
        assert np.array_equal(ids, np.flatnonzero((layer >= z0) & (layer < z1)))
        assert np.array_equal(hp["position"][:, :3].view(np.uint32), o.pos[ids].view(np.uint32))
        assert sim.max_count == o.n  # the scene size stays the global one
        seen[ids] += 1
    assert np.all(seen == 1)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch, torch.distributed as dist
import gmu_water_simulation_b200 as gws
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
box = (0.4, 0.4, 0.9)
ident = [bytes(range(128)) if rank == 0 else None]       # stands in for the NCCL id: same plumbing as bench.py
dist.broadcast_object_list(ident, src=0)
assert ident[0] == bytes(range(128))
rz = int(gws.make_config(box, 1).grid_res[2])
z0, z1 = gws.slab_plan(rz, world, rank)
sim = gws.Simulator("scene_only", box).set_owned_layers(z0, z1).setup_scene()
ids = sim.host_particles()["id"].astype(np.int64)
t = torch.tensor([sim.n, int(ids.sum()), int(ids.min()), int(ids.max())], dtype=torch.int64)
parts = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(parts, t)
total = sum(int(p[0]) for p in parts)
n_all = sim.max_count
assert total == n_all, (total, n_all)
assert sum(int(p[1]) for p in parts) == n_all * (n_all - 1) // 2      # every id exactly once
agg = torch.tensor([float(sim.n)], dtype=torch.float64)
dist.all_reduce(agg, op=dist.ReduceOp.SUM)                             # the bench's whole-job aggregation
assert int(agg.item()) == n_all
if rank == 0:
    print("GLOO SLAB OK", total)
dist.destroy_process_group()
'''


def test_world2_gloo_scene_partition(tmp_path):
    """N>1 host path on CPU: rendezvous + id broadcast + slab plan + per-rank scene generation + aggregation."""
    worker = tmp_path / "worker.py"
    worker.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(worker), ROOT]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO SLAB OK" in out.stdout


@pytest.mark.gpu
def test_single_rank_slab_equals_plain_run(gws):
    """world == 1 slab mode (no neighbours) runs the slab step loop; results must equal the plain path bitwise."""
    box = (0.4, 0.4, 0.9)
    plain = gws.Simulator("cuda", box).setup_scene()
    slab = gws.Simulator("cuda", box).enable_slab(0, 1, bytes(128)).setup_scene()
    plain.step_many(12)
    slab.step_many(12)
    plain.sync_host()
    a = plain.host_particles()
    b = slab.context().download_owned()
    order = np.argsort(b["id"], kind="stable")
    b = b[order]
    assert np.array_equal(b["id"], a["id"])
    for f in ("position", "velocity", "density", "cell_id"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
    info = slab.context().slab_info()
    assert info["n_own"] == plain.n and info["z0"] == 0
    assert slab.context().counter("slab_far_movers") == 0
    # the sequential (non-overlapped) slab step gives the same bits
    seq = gws.Simulator("cuda", box).enable_slab(0, 1, bytes(128)).setup_scene()
    seq.context().set_option("slab_overlap", 0)
    seq.step_many(12)
    c = seq.context().download_owned()
    c = c[np.argsort(c["id"], kind="stable")]
    assert np.array_equal(b["position"].view(np.uint32), c["position"].view(np.uint32))


def run_loopback_slabs(gws, box, world, steps, gravity=None, device_of_rank=None, overlap=True, rebalance=None):
    """`world` slab ranks inside this process over the loop-back transport (sph_comm_local_id), each stepped from
    its own host thread like a rank process would; returns the merged owned particles sorted by id + the contexts' info."""
    import threading

    ident = gws.comm_local_id(world)
    sims = []
    for rank in range(world):
        dev = device_of_rank(rank) if device_of_rank else 0
        sim = gws.Simulator("cuda", box, device=dev).enable_slab(rank, world, ident).setup_scene()
        if gravity is not None:
            sim.set_gravity(gravity)
        if not overlap:
            sim.context().set_option("slab_overlap", 0)
        if rebalance is not None:
            sim.context().set_option("slab_rebalance", rebalance)
        sims.append(sim)
    errors = []

    def work(sim):
        try:
            sim.step_many(steps)
        except Exception as exc:  # noqa: BLE001 - reported by the main thread
            errors.append(exc)
            sim.close()  # destroying the rank wakes the neighbours that wait for it (they fail with a COMM error)

    threads = [threading.Thread(target=work, args=(sim,), daemon=True) for sim in sims]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a slab rank is stuck in the exchange"
    assert not errors, errors
    parts = [sim.context().download_owned() for sim in sims]
    infos = [sim.context().slab_info() for sim in sims]
    far = [sim.context().counter("slab_far_movers") for sim in sims]
    for info, sim in zip(infos, sims):
        info["face_moves"] = sim.context().counter("slab_face_moves")
    merged = np.concatenate(parts)
    merged = merged[np.argsort(merged["id"], kind="stable")]
    for sim in sims:
        sim.close()
    return merged, parts, infos, far


def migrated_particles(gws, box, world, parts):
    """How many particles are now owned by another rank than the one whose layers they started in."""
    rz = int(gws.make_config(box, 1).grid_res[2])
    scene = gws.Simulator("scene_only", box).setup_scene()  # keep the owner of the host mirror alive while it is read
    z = scene.host_particles()["position"][:, 2].astype(np.float64)
    layer0 = np.clip(np.floor((z + np.float32(box[2]) / 2.0) / np.float64(np.float32(0.0457))), 0, rz - 1)
    scene.close()
    moved = 0
    for rank, part in enumerate(parts):
        z0, z1 = gws.slab_plan(rz, world, rank)
        l0 = layer0[part["id"]]
        moved += int(((l0 < z0) | (l0 >= z1)).sum())
    return moved


def assert_same_as_plain(gws, merged, box, steps, gravity=None):
    plain = gws.Simulator("cuda", box).setup_scene()
    if gravity is not None:
        plain.set_gravity(gravity)
    plain.step_many(steps)
    plain.sync_host()
    hp = plain.host_particles()
    assert merged.shape[0] == plain.n, "the slabs do not hold every particle exactly once"
    assert np.array_equal(merged["id"], np.arange(plain.n, dtype=np.uint32)), "owned sets are not a partition of the ids"
    for f in ("cell_id", "grid_position", "position", "velocity", "density", "pressure", "acceleration"):
        same = merged[f].view(np.uint32) == hp[f].view(np.uint32)
        assert same.all(), f"{f}: {(~same).sum()} words differ from the single-GPU run"
    plain.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("overlap", [True, False])
def test_loopback_slabs_equal_one_gpu_small_tank_with_migration(gws, world, overlap):
    """k slabs over the loop-back transport == one GPU, BIT FOR BIT, on a small tank whose water is pushed along z
    (gravity tilted towards +z) so that particles keep migrating across the slab faces for 60 steps."""
    box, steps, g = (0.5, 0.5, 1.7), 60, (0.0, -9.80665, 6.0)
    merged, parts, infos, far = run_loopback_slabs(gws, box, world, steps, gravity=g, overlap=overlap)
    assert far == [0] * world
    assert migrated_particles(gws, box, world, parts) > 50  # the ranks no longer own the particles they started with
    assert_same_as_plain(gws, merged, box, steps, gravity=g)


@pytest.mark.gpu
@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("rebalance", [2, 1, 0])
def test_loopback_slabs_with_moving_faces_equal_one_gpu(gws, overlap, rebalance):
    """The faces between slabs move while the run goes on (load balance): mode 2 drives every face up and down by a
    deterministic pattern, mode 1 by the measured loads, mode 0 keeps them fixed.  Whatever the faces do, the merged
    state is bit-identical to the single-GPU run, every particle is owned exactly once, and no particle is missed."""
    box, steps, g, world = (0.5, 0.5, 2.6), 40, (0.0, -9.80665, 5.0), 3   # 57 cell layers: 19 per slab
    merged, parts, infos, far = run_loopback_slabs(gws, box, world, steps, gravity=g, overlap=overlap, rebalance=rebalance)
    assert far == [0] * world
    moves = sum(i["face_moves"] for i in infos)
    if rebalance == 2:
        assert moves >= 2 * 20, moves                      # both faces moved on most steps (counted on both sides)
    if rebalance == 0:
        assert moves == 0 and [(i["z0"], i["z1"]) for i in infos] == [(0, 19), (19, 38), (38, 57)]
    assert [i["z1"] for i in infos[:-1]] == [i["z0"] for i in infos[1:]] and infos[0]["z0"] == 0 and infos[-1]["z1"] == 57
    assert_same_as_plain(gws, merged, box, steps, gravity=g)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_loopback_slabs_equal_one_gpu_4m_tank(gws, world):
    """SURVEY.md §8e equivalence case: the 4,000,000-particle tank 4.56 x 4.56 x 9.13 (100 x 100 x 200 cells), k = 2
    and 4 slabs vs one GPU after 14 steps, bit for bit (keys, positions, velocities, density, pressure, acceleration).
    Gravity is tilted towards +z so that particles cross the slab faces (migration) during the run."""
    box, steps, g = (4.56, 4.56, 9.13), 14, (0.0, -9.80665, 6.0)
    merged, parts, infos, far = run_loopback_slabs(gws, box, world, steps, gravity=g)
    assert merged.shape[0] == 4_000_000 and far == [0] * world
    assert sum(i["z1"] - i["z0"] for i in infos) == 200 and infos[0]["z0"] == 0 and infos[-1]["z1"] == 200
    assert migrated_particles(gws, box, world, parts) > 1000
    assert_same_as_plain(gws, merged, box, steps, gravity=g)


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_nccl_slab_equivalence(gws, nproc):
    """`nproc` ranks on `nproc` GPUs over NCCL vs one GPU (tools/slab_check.py); skipped where the box has fewer GPUs."""
    if gws.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    box = ["--box", "0.5", "0.5", str(0.0457 * 4.2 * nproc + 0.9)]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + nproc), os.path.join(ROOT, "tools", "slab_check.py"), "--steps", "20", "--bitwise"] + box
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SLAB CHECK OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


@pytest.mark.gpu
def test_two_rank_slab_equivalence(gws):
    """2 ranks on 2 GPUs vs one GPU (tools/slab_check.py); skipped on single-GPU boxes."""
    if gws.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for extra in ([], ["--no-overlap"]):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29534", os.path.join(ROOT, "tools", "slab_check.py"), "--steps", "20"] + extra
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "SLAB CHECK OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])
