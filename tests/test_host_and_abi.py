"""CPU-only checks: the C-ABI libraries load and export every declared symbol; host logic (scene
generation, emission, error behaviour without a device) of the C++ simulator core."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_binding import FOUNTAIN, Oracle


def test_cuda_library_exports_every_declared_symbol(gws):
    lib = gws.cuda_lib()
    names = gws.declared_symbols("sph_cuda.h")
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"libsph_cuda.so does not export {name}"
    assert lib.sph_abi_version() == 1


def test_host_library_exports_every_declared_symbol(gws):
    lib = gws.host_lib()
    for name in [n for n in gws.declared_symbols("sph_host.h") if n.startswith("gmu_sim_")]:
        assert hasattr(lib, name), f"libsph_host.so does not export {name}"


def test_abi_record_sizes(gws):
    assert gws.PARTICLE_DTYPE.itemsize == 80
    cfg = gws.make_config(0.9, 16000)
    assert tuple(cfg.grid_res) == (20, 20, 20)
    assert cfg.wall_count == 6
    assert [w.normal[i % 3] for i, w in enumerate(cfg.walls)] == [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]
    assert cfg.walls[0].position[0] == -np.float32(0.9) / 2 and cfg.walls[4].position[1] == np.float32(0.9) / 2
    assert abs(cfg.h - 0.0457) < 1e-7 and abs(cfg.gravity[1] + 9.80665) < 1e-6


def test_no_cpu_fallback_without_device(gws):
    """On a box without a GPU every compute entry point must fail loudly, not fall back."""
    try:
        n = gws.device_count()
    except gws.SphError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gws.SphError):
        gws.SphContext(0.4, 1620)
    sim = gws.Simulator("cuda", 0.4)
    with pytest.raises(gws.SphError):
        sim.setup_scene()


@pytest.mark.parametrize("box", [0.4, 0.9, (0.5, 0.3, 0.7)])
def test_host_scene_generator_matches_oracle_bitwise(gws, box):
    sim = gws.Simulator("scene_only", box).setup_scene()
    o = Oracle(box).setup_scene()
    hp = sim.host_particles()
    assert sim.n == o.n == sim.max_count
    assert np.array_equal(hp["position"][:, :3].view(np.uint32), o.pos.view(np.uint32))
    assert np.array_equal(hp["id"], np.arange(o.n, dtype=np.uint32))
    assert not hp["velocity"].any() and not hp["density"].any()


def test_host_fountain_emitter_matches_oracle(gws):
    sim = gws.Simulator("scene_only", 0.4, scenario=gws.FOUNTAIN).setup_scene()
    o = Oracle(0.4, FOUNTAIN).setup_scene()
    assert sim.n == 0 and sim.max_count == o.max_count == 1620
    sim.emit(300)
    for _ in range(300):
        o.generate_particles()
    assert sim.n == o.n == 1617 and sim.iteration == 300
    hp = sim.host_particles()
    assert np.array_equal(hp["position"][:, :3].view(np.uint32), o.pos.view(np.uint32))
    assert np.array_equal(hp["velocity"][:, :3].view(np.uint32), o.vel.view(np.uint32))


def test_emission_multiplier_nozzles_do_not_coincide(gws):
    sim = gws.Simulator("scene_only", 2.28, scenario=gws.FOUNTAIN).setup_scene()
    sim.set_emission_multiplier(9)
    sim.emit(1)
    assert sim.n == 63
    p = sim.host_particles()["position"][:, :3]
    d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2) + np.eye(63)
    assert d.min() > 0.01


def test_gravity_keys_follow_reference(gws):
    """onKeyPressed / toggleGravity / toggleSimulation (src/CBaseParticleSimulator.cpp:84-114,154-179): G toggles
    gravity between (0, g, 0) and 0 — a tilt is lost by the toggle —, O / P tilt x by -/+1, space pauses/resumes.
    Where oracle/_ref is built the same key sequence is replayed on the reference's own object and compared."""
    import ref_binding

    sim = gws.Simulator("scene_only", 0.4).setup_scene()
    ref = ref_binding.Reference(0.4).setup_scene() if ref_binding.available() else None
    g0 = np.array([0.0, -9.80665, 0.0], dtype=np.float32)
    expected = {0: g0}
    seq = [(0x4F, g0 + np.float32([-1, 0, 0])), (0x50, g0), (0x50, g0 + np.float32([1, 0, 0])), (0x50, g0 + np.float32([2, 0, 0])),
           (0x47, np.zeros(3, np.float32)), (0x4F, np.float32([-1, 0, 0])), (0x47, np.zeros(3, np.float32)), (0x47, g0)]
    assert np.array_equal(sim.gravity, g0)
    for key, want in seq:
        sim.key(key)
        assert np.array_equal(sim.gravity, want), (hex(key), sim.gravity, want)
        if ref is not None:
            ref.key(key)
            assert np.array_equal(sim.gravity.view(np.uint32), ref.gravity.view(np.uint32)), hex(key)
    assert not sim.running
    sim.key(0x20)
    assert sim.running
    sim.key(0x20)
    assert not sim.running
    it = sim.iteration
    sim.key(0x53)  # S: one doWork()
    assert sim.iteration == it + 1


def test_simulation_type_values_keep_the_reference_rows(gws):
    """include/mainwindow.h:60-63: GPUGrid = 0, GPUBrute, CPU are used as combo-box row indices
    (src/mainwindow.cpp:90-97); the CUDA types are appended."""
    assert [gws.simulation_type(t) for t in ("GPU Grid", "GPU Brute Force", "CPU Grid", "CUDA Grid", "CUDA Brute Force")] == [0, 1, 2, 3, 4]
    assert gws.simulation_type("OpenGL") == -1
    for t in (0, 1, 2):  # the OpenCL / CPU simulators are not built here: the factory refuses, it does not fall back
        with pytest.raises(gws.SphError, match="only the CUDA simulators"):
            gws.Simulator(t, 0.4)
    sim = gws.Simulator(3, 0.4)  # constructing needs no device; setupScene does
    assert sim.max_count == 0


def test_unknown_simulator_type(gws):
    with pytest.raises(gws.SphError):
        gws.Simulator("opencl", 0.4)


def test_export_logs_match_the_reference_layout_byte_for_byte(gws, tmp_path):
    """MainWindow::exportLogs (src/mainwindow.cpp:310-368): file names "<Scenario text>_<box>" with the combo
    text "Dam break" (:103), every value followed by ';' (:332-339), the detail block = name line, five labelled
    rows, two blank lines (:350-362), append mode (:323), numbers as QString::number = %g."""
    sim = gws.Simulator("scene_only", 0.4).setup_scene()
    sim.push_event(0, 60.0, 0.5, 1.25, 2.0, 0.0, 0.125)
    sim.push_event(10, 61.0, 1.0, 2.5, 4.0, 0.0, 0.25)
    sim.export_logs(str(tmp_path), "CUDA Grid")
    data = (tmp_path / "Dam break_0.4.csv").read_bytes()
    detail = (tmp_path / "Dam break_0.4_detail.csv").read_bytes()
    assert data == b"CUDA Grid;3.875;7.75;\n"
    block = (b"CUDA Grid\n"
             b"Grid;0.5;1;\n"
             b"Density + pressure;1.25;2.5;\n"
             b"Forces;2;4;\n"
             b"Collisions;0;0;\n"
             b"Integrate;0.125;0.25;\n"
             b"\n\n")
    assert detail == block
    sim.export_logs(str(tmp_path), "CUDA Grid")  # a second run appends
    assert (tmp_path / "Dam break_0.4.csv").read_bytes() == 2 * data
    assert (tmp_path / "Dam break_0.4_detail.csv").read_bytes() == 2 * block
    # QString::number(double) is %g: 6 significant digits, exponent form for small values
    f = gws.Simulator("scene_only", 1.0, scenario=gws.FOUNTAIN).setup_scene()
    f.push_event(0, 0.0, 1.0 / 3.0, 0.00001234567, 123456.789, 0.0, 0.0)
    f.export_logs(str(tmp_path), "CUDA Grid")
    rows = (tmp_path / "Fountain_1_detail.csv").read_bytes().split(b"\n")
    assert rows[1] == b"Grid;0.333333;" and rows[2] == b"Density + pressure;1.23457e-05;" and rows[3] == b"Forces;123457;"


def test_export_logs_from_profiled_steps(gws, tmp_path):
    """The sampled-step path: every eventLoggerStride-th step appends one record (src/CBaseParticleSimulator.cpp:138-143)."""
    sim = gws.Simulator("scene_only", 0.4).setup_scene()
    sim.set_profiling(True, 2)
    sim.emit(5)  # iterations 0..4 -> samples at 0, 2, 4
    sim.export_logs(str(tmp_path), "CUDA Grid")
    fields = (tmp_path / "Dam break_0.4.csv").read_text().rstrip("\n").split(";")
    assert fields[0] == "CUDA Grid" and fields[-1] == "" and len(fields) == 5


@pytest.mark.parametrize("header", ["sph_cuda.h", "sph_host.h"])
def test_public_headers_are_plain_c(header, tmp_path):
    """The boundary is a C ABI: both headers must compile as C99 on their own (no C++, no CUDA or torch types)."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "t.c"
    src.write_text(f'#include "{header}"\nint main(void) {{ return 0; }}\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(src)])


def _build_minimal_step(gws, tmp_path):
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(gws.binding._CUDA_SO)
    exe = str(tmp_path / "minimal_step")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "minimal_step.c"),
                           "-L", libdir, "-lsph_cuda", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe])
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_plain_c_example_links_and_fails_loudly_without_a_device(gws, tmp_path):
    """examples/minimal_step.c is the ABI used from C.  It must link against libsph_cuda.so alone; without a GPU it
    reports the CUDA error and exits 1 (no CPU fallback)."""
    try:
        has_gpu = gws.device_count() > 0
    except gws.SphError:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present: covered by test_plain_c_example_steps_on_the_gpu")
    run = _build_minimal_step(gws, tmp_path)
    assert run.returncode == 1 and "CUDA::" in run.stderr


@pytest.mark.gpu
def test_plain_c_example_steps_on_the_gpu(gws, tmp_path):
    """The same C program on a B200: 16 000 particles stepped through the C ABI from plain C."""
    run = _build_minimal_step(gws, tmp_path)
    assert run.returncode == 0 and "16000 particles" in run.stdout, run.stderr
