"""The bench.py contract that can be checked without a GPU: the reference arm's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--reference-sample", "dam_break_16K"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "dam_break_1M" and d["config"]["sample"] == "dam_break_16K"
    import ref_binding

    # the reference's own code (oracle/_ref) wherever it is built; the bit-identical port only where it is not
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_binding.available() else "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 1e5 < d["value"] < 1e7  # a single host core does a few 1e5 particle-steps/s


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import gmu_water_simulation_b200 as gws

    try:
        if gws.device_count() > 0:
            return
    except gws.SphError:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_ncu_facts_belong_to_the_kernels_in_the_tree(gws):
    """profiles/kernel_traffic.json (dram bytes and pipe utilisations bench.py reports as roofline.traffic /
    binding_roof) must come from an ncu capture of the kernel sources that are in the tree; bench.py refuses the file
    otherwise.  Re-run tools/make_profiles.sh + tools/kernel_traffic.py after changing a kernel."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    facts = bench.ncu_facts(gws)
    assert facts is not None, "profiles/kernel_traffic.json is stale: it was captured from other kernel sources"
    for kernel in ("k_density_mask", "k_forces_mask"):
        k = facts["kernels"][kernel]
        assert k["dram_bytes"] > 1e7 and 0 < k["issue_active_frac"] < 1 and 0 < k["l1_data_pipe_frac"] < 1
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    stale = dict(facts, kernel_source_sha256="0" * 64)
    backup = open(path).read()
    try:
        with open(path, "w") as f:
            json.dump(stale, f)
        assert bench.ncu_facts(gws) is None
    finally:
        with open(path, "w") as f:
            f.write(backup)
