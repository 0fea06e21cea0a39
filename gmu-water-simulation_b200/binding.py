"""ctypes bindings of include/sph_cuda.h (C ABI) and include/sph_host.h (simulator facade)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# SPH_CUDA_SO: development switch for A/B runs of two builds of the kernel library (tools/kernel_probe.py)
_CUDA_SO = os.environ.get("SPH_CUDA_SO") or os.path.join(_PKG, "libsph_cuda.so")
_HOST_SO = os.path.join(_PKG, "libsph_host.so")

DAM_BREAK, FOUNTAIN = 0, 1

# CParticle::Physics / sph_particle: 80 bytes (include/sph_cuda.h)
PARTICLE_DTYPE = np.dtype(
    [
        ("position", np.float32, 4),
        ("velocity", np.float32, 4),
        ("acceleration", np.float32, 4),
        ("grid_position", np.int32, 4),
        ("density", np.float32),
        ("pressure", np.float32),
        ("id", np.uint32),
        ("cell_id", np.uint32),
    ],
    align=False,
)
assert PARTICLE_DTYPE.itemsize == 80


class SphError(RuntimeError):
    pass


class _Wall(C.Structure):
    _fields_ = [("normal", C.c_float * 4), ("position", C.c_float * 4)]


class SphConfig(C.Structure):
    _fields_ = [
        ("box", C.c_float * 3),
        ("grid_res", C.c_int32 * 3),
        ("h", C.c_float),
        ("dt", C.c_float),
        ("mass", C.c_float),
        ("viscosity", C.c_float),
        ("gas_stiffness", C.c_float),
        ("rest_density", C.c_float),
        ("gravity", C.c_float * 3),
        ("wall_k", C.c_double),
        ("wall_damping", C.c_double),
        ("wall_skin", C.c_double),
        ("wall_count", C.c_int32),
        ("walls", _Wall * 6),
        ("max_particles", C.c_uint32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("nccl_id", C.c_uint8 * 128),
    ]


def build(verbose=False):
    """Compile libsph_cuda.so (nvcc, sm_100a) and libsph_host.so (g++) in-tree."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_PKG, "csrc")], stdout=out)
    subprocess.check_call(["make", "-C", os.path.join(_PKG, "core")], stdout=out)
    return _CUDA_SO, _HOST_SO


def kernel_source_hash():
    """sha256 over the sources of the two neighbour kernels an ncu capture describes (csrc/sph_neighbours_v2.cu, the shared
    device header, the launcher header, the Makefile with the compiler flags): identifies the build a capture belongs
    to.  (The .so itself is not byte-reproducible across nvcc runs.)"""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(_PKG, "csrc")
    for name in ("sph_neighbours_v2.cu", "sph_device.cuh", "sph_kernels.h", "Makefile"):
        h.update(name.encode() + b"\0")
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def declared_symbols(header):
    """Function names declared in include/<header> (used by the export test)."""
    text = open(os.path.join(_ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sph|gmu_sim)_[a-z0-9_]+)\s*\(", text)))


_cuda = None
_host = None


def cuda_lib():
    """Load libsph_cuda.so; fails loudly when it has not been built (no CPU fallback)."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(_CUDA_SO):
            raise SphError(f"{_CUDA_SO} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(_CUDA_SO, mode=C.RTLD_GLOBAL)
        vp, i32, u32, u64, dbl, f = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double, C.c_float
        P = C.POINTER
        sig = {
            "sph_abi_version": (i32, []),
            "sph_device_count": (i32, [P(i32)]),
            "sph_device_name": (i32, [i32, C.c_char_p, C.c_size_t]),
            "sph_config_init": (i32, [P(SphConfig), f, f, f, u32]),
            "sph_create": (i32, [P(SphConfig), P(vp)]),
            "sph_destroy": (i32, [vp]),
            "sph_last_error": (C.c_char_p, [vp]),
            "sph_upload_particles": (i32, [vp, vp, u32]),
            "sph_append_particles": (i32, [vp, vp, u32]),
            "sph_state_save": (i32, [vp, i32]),
            "sph_state_restore": (i32, [vp, i32]),
            "sph_set_emitter": (i32, [vp, vp, u32, u32, u32]),
            "sph_emit": (i32, [vp, P(u32)]),
            "sph_download_particles": (i32, [vp, vp, u32, P(u32)]),
            "sph_download_particles_async": (i32, [vp, vp, u32]),
            "sph_download_wait": (i32, [vp, P(u32)]),
            "sph_particle_count": (i32, [vp, P(u32)]),
            "sph_set_gravity": (i32, [vp, f, f, f]),
            "sph_set_collision_faces": (i32, [vp, vp, u32]),
            "sph_pin_host_buffer": (i32, [vp, vp, C.c_size_t]),
            "sph_update_grid": (i32, [vp, P(dbl)]),
            "sph_density_pressure": (i32, [vp, P(dbl)]),
            "sph_forces": (i32, [vp, P(dbl)]),
            "sph_collisions": (i32, [vp, P(dbl)]),
            "sph_integrate": (i32, [vp, P(dbl)]),
            "sph_step": (i32, [vp, i32, P(dbl)]),
            "sph_step_profiled": (i32, [vp, i32, vp]),
            "sph_synchronize": (i32, [vp]),
            "sph_download_keys": (i32, [vp, vp]),
            "sph_download_permutation": (i32, [vp, vp]),
            "sph_download_cell_start": (i32, [vp, vp]),
            "sph_download_density_pressure_accel": (i32, [vp, vp, vp, vp]),
            "sph_download_neighbours": (i32, [vp, vp, vp, u64, P(u64)]),
            "sph_download_mask_neighbours": (i32, [vp, vp, vp, u64, P(u64)]),
            "sph_brute_density_pressure": (i32, [vp, P(dbl)]),
            "sph_brute_forces": (i32, [vp, P(dbl)]),
            "sph_brute_neighbour_counts": (i32, [vp, vp]),
            "sph_stats": (i32, [vp, vp]),
            "sph_fill_height_percentile": (i32, [vp, dbl, P(dbl)]),
            "sph_set_option": (i32, [vp, C.c_char_p, i32]),
            "sph_get_counter": (i32, [vp, C.c_char_p, P(u64)]),
            "sph_comm_unique_id": (i32, [vp]),
            "sph_comm_local_id": (i32, [i32, vp]),
            "sph_slab_plan": (i32, [i32, i32, i32, P(i32), P(i32)]),
            "sph_slab_create": (i32, [P(SphConfig), P(vp)]),
            "sph_slab_face_shift": (i32, [vp, vp, i32, i32, i32, u64, i32]),
            "sph_slab_info": (i32, [vp, vp]),
            "sph_download_owned": (i32, [vp, vp, u32, P(u32)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _cuda = lib
    return _cuda


def host_lib():
    """Load libsph_host.so (the C++ simulator core)."""
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(_HOST_SO):
            raise SphError(f"{_HOST_SO} is missing: run __graft_entry__.build().")
        lib = C.CDLL(_HOST_SO)
        vp, i32, u64, dbl, f = C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_float
        sig = {
            "gmu_sim_create": (vp, [C.c_char_p, f, f, f, i32, i32]),
            "gmu_sim_destroy": (None, [vp]),
            "gmu_sim_last_error": (C.c_char_p, []),
            "gmu_sim_enable_slab": (i32, [vp, i32, i32, vp]),
            "gmu_sim_set_owned_layers": (i32, [vp, i32, i32]),
            "gmu_sim_setup_scene": (i32, [vp]),
            "gmu_sim_step": (i32, [vp, i32]),
            "gmu_sim_step_many": (i32, [vp, i32, C.POINTER(dbl)]),
            "gmu_sim_emit": (i32, [vp, i32]),
            "gmu_sim_set_mirror_mode": (i32, [vp, i32]),
            "gmu_sim_set_mirror_stride": (i32, [vp, i32]),
            "gmu_sim_sync_host": (i32, [vp]),
            "gmu_sim_wait_host": (i32, [vp]),
            "gmu_sim_set_gravity": (i32, [vp, f, f, f]),
            "gmu_sim_set_collision_faces": (i32, [vp, vp, i32]),
            "gmu_sim_key": (i32, [vp, i32]),
            "gmu_sim_get_gravity": (i32, [vp, vp]),
            "gmu_sim_is_running": (i32, [vp]),
            "gmu_sim_push_event": (i32, [vp, vp]),
            "gmu_sim_type_from_name": (i32, [C.c_char_p]),
            "gmu_sim_create_by_type": (vp, [i32, f, f, f, i32, i32]),
            "gmu_sim_set_profiling": (i32, [vp, i32, i32]),
            "gmu_sim_set_emission_multiplier": (i32, [vp, i32]),
            "gmu_sim_particle_count": (u64, [vp]),
            "gmu_sim_max_particle_count": (u64, [vp]),
            "gmu_sim_iteration": (u64, [vp]),
            "gmu_sim_host_particles": (vp, [vp]),
            "gmu_sim_context": (vp, [vp]),
            "gmu_sim_device_name": (C.c_char_p, [vp]),
            "gmu_sim_event_count": (u64, [vp]),
            "gmu_sim_get_events": (u64, [vp, vp, u64]),
            "gmu_sim_export_logs": (i32, [vp, C.c_char_p, C.c_char_p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _host = lib
    return _host


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def simulation_type(combo_text):
    """eSimulationType value of a simulation-type combo text ("GPU Grid" -> 0 ... "CUDA Grid" -> 3); -1 if unknown."""
    return int(host_lib().gmu_sim_type_from_name(combo_text.encode()))


def device_count():
    n = C.c_int(0)
    rc = cuda_lib().sph_device_count(C.byref(n))
    if rc:
        raise SphError(cuda_lib().sph_last_error(None).decode())
    return n.value


def device_name(device=0):
    buf = C.create_string_buffer(256)
    rc = cuda_lib().sph_device_name(device, buf, 256)
    if rc:
        raise SphError(cuda_lib().sph_last_error(None).decode())
    return buf.value.decode()


def comm_unique_id():
    """128-byte NCCL id for slab mode (create on rank 0, broadcast to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    rc = cuda_lib().sph_comm_unique_id(buf)
    if rc:
        raise SphError(cuda_lib().sph_last_error(None).decode())
    return bytes(buf)


def comm_local_id(world):
    """128-byte id of the loop-back transport: `world` slab ranks inside this process, one host thread each."""
    buf = (C.c_uint8 * 128)()
    rc = cuda_lib().sph_comm_local_id(int(world), buf)
    if rc:
        raise SphError(cuda_lib().sph_last_error(None).decode())
    return bytes(buf)


def slab_face_shift(lower, upper, shift=0, shift_max=8, mode=1, exchange=0, face=0):
    """The rule both ranks at a slab face apply to the records {load_us, z0, z1, may_grow} of the lower / upper rank."""
    lo = np.asarray(lower, dtype=np.int32)
    hi = np.asarray(upper, dtype=np.int32)
    return int(cuda_lib().sph_slab_face_shift(_ptr(lo), _ptr(hi), int(shift), int(shift_max), int(mode), int(exchange), int(face)))


def slab_plan(rz, world, rank):
    """Owned z-layers [z0, z1) of `rank` (pure host logic, no device needed)."""
    z0, z1 = C.c_int32(0), C.c_int32(0)
    rc = cuda_lib().sph_slab_plan(int(rz), int(world), int(rank), C.byref(z0), C.byref(z1))
    if rc:
        raise SphError(cuda_lib().sph_last_error(None).decode())
    return z0.value, z1.value


def make_config(box, max_particles, device=0):
    if np.isscalar(box):
        box = (box, box, box)
    cfg = SphConfig()
    cuda_lib().sph_config_init(C.byref(cfg), float(box[0]), float(box[1]), float(box[2]), int(max_particles))
    cfg.device = device
    return cfg


def particles_from_arrays(pos, vel=None, ids=None):
    """Build the 80-byte AoS records from (n,3) arrays."""
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    n = pos.shape[0]
    rec = np.zeros(n, dtype=PARTICLE_DTYPE)
    rec["position"][:, :3] = pos
    if vel is not None:
        rec["velocity"][:, :3] = np.asarray(vel, dtype=np.float32).reshape(-1, 3)
    rec["id"] = np.arange(n, dtype=np.uint32) if ids is None else np.asarray(ids, dtype=np.uint32)
    return rec


class SphContext:
    """One device context of the C ABI (≙ one CLWrapper + its buffers)."""

    def __init__(self, box, max_particles, device=0, cfg=None):
        self.lib = cuda_lib()
        self.cfg = cfg if cfg is not None else make_config(box, max_particles, device)
        h = C.c_void_p()
        rc = self.lib.sph_create(C.byref(self.cfg), C.byref(h))
        if rc:
            raise SphError(self.lib.sph_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sph_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise SphError(f"[{rc}] " + self.lib.sph_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    @property
    def grid_res(self):
        return tuple(self.cfg.grid_res)

    @property
    def n_cells(self):
        r = self.grid_res
        return r[0] * r[1] * r[2]

    @property
    def n(self):
        v = C.c_uint32(0)
        self._ck(self.lib.sph_particle_count(self._h, C.byref(v)))
        return v.value

    # state
    def upload(self, rec):
        rec = np.ascontiguousarray(rec, dtype=PARTICLE_DTYPE)
        self._ck(self.lib.sph_upload_particles(self._h, _ptr(rec), rec.shape[0]))

    def append(self, rec):
        rec = np.ascontiguousarray(rec, dtype=PARTICLE_DTYPE)
        self._ck(self.lib.sph_append_particles(self._h, _ptr(rec), rec.shape[0]))

    def state_save(self, slot=0):
        self._ck(self.lib.sph_state_save(self._h, int(slot)))

    def state_restore(self, slot=0):
        self._ck(self.lib.sph_state_restore(self._h, int(slot)))

    def set_emitter(self, templates, group=7, max_count=None):
        rec = np.ascontiguousarray(templates, dtype=PARTICLE_DTYPE)
        self._ck(self.lib.sph_set_emitter(self._h, _ptr(rec) if rec.shape[0] else None, rec.shape[0], int(group),
                                          int(self.cfg.max_particles if max_count is None else max_count)))

    def emit(self):
        got = C.c_uint32(0)
        self._ck(self.lib.sph_emit(self._h, C.byref(got)))
        return got.value

    def download(self, out=None):
        n = self.n
        if out is None:
            out = np.zeros(n, dtype=PARTICLE_DTYPE)
        got = C.c_uint32(0)
        self._ck(self.lib.sph_download_particles(self._h, _ptr(out), out.shape[0], C.byref(got)))
        return out[: got.value]

    def download_async(self, out):
        """Start an overlapped read-back into `out` (pin it first); contents are defined after download_wait()."""
        self._ck(self.lib.sph_download_particles_async(self._h, _ptr(out), out.shape[0]))

    def download_wait(self):
        got = C.c_uint32(0)
        self._ck(self.lib.sph_download_wait(self._h, C.byref(got)))
        return got.value

    def pin(self, arr):
        self._ck(self.lib.sph_pin_host_buffer(self._h, _ptr(arr), arr.nbytes))

    def set_gravity(self, g):
        self._ck(self.lib.sph_set_gravity(self._h, float(g[0]), float(g[1]), float(g[2])))

    def set_collision_faces(self, faces):
        """faces: (n, 12) float32 = normal, v0, v1, v2 per face (sph_face); an empty array clears the mesh."""
        faces = np.ascontiguousarray(faces, dtype=np.float32).reshape(-1, 12)
        self._ck(self.lib.sph_set_collision_faces(self._h, _ptr(faces) if faces.shape[0] else None, faces.shape[0]))

    # phases: return device ms when timed=True
    def _phase(self, fn, timed):
        if timed:
            ms = C.c_double(0)
            self._ck(fn(self._h, C.byref(ms)))
            return ms.value
        self._ck(fn(self._h, None))
        return 0.0

    def update_grid(self, timed=True):
        return self._phase(self.lib.sph_update_grid, timed)

    def density_pressure(self, timed=True):
        return self._phase(self.lib.sph_density_pressure, timed)

    def forces(self, timed=True):
        return self._phase(self.lib.sph_forces, timed)

    def collisions(self, timed=True):
        return self._phase(self.lib.sph_collisions, timed)

    def integrate(self, timed=True):
        return self._phase(self.lib.sph_integrate, timed)

    def brute_density_pressure(self, timed=True):
        return self._phase(self.lib.sph_brute_density_pressure, timed)

    def brute_forces(self, timed=True):
        return self._phase(self.lib.sph_brute_forces, timed)

    def step(self, n=1, timed=True):
        if timed:
            ms = C.c_double(0)
            self._ck(self.lib.sph_step(self._h, int(n), C.byref(ms)))
            return ms.value
        self._ck(self.lib.sph_step(self._h, int(n), None))
        return 0.0

    def step_profiled(self, n=1):
        """n steps with events between the kernel groups: dict of summed ms (grid, density, forces, total)."""
        out = np.zeros(4, dtype=np.float64)
        self._ck(self.lib.sph_step_profiled(self._h, int(n), _ptr(out)))
        return dict(grid=out[0], density=out[1], forces=out[2], total=out[3])

    def synchronize(self):
        self._ck(self.lib.sph_synchronize(self._h))

    # taps
    def keys(self):
        out = np.empty(self.n, dtype=np.int32)
        self._ck(self.lib.sph_download_keys(self._h, _ptr(out)))
        return out

    def permutation(self):
        out = np.empty(self.n, dtype=np.uint32)
        self._ck(self.lib.sph_download_permutation(self._h, _ptr(out)))
        return out

    def cell_start(self):
        out = np.empty(self.n_cells + 1, dtype=np.int32)
        self._ck(self.lib.sph_download_cell_start(self._h, _ptr(out)))
        return out

    def density_pressure_accel(self):
        n = self.n
        rho, prs, acc = np.empty(n, np.float32), np.empty(n, np.float32), np.empty((n, 3), np.float32)
        self._ck(self.lib.sph_download_density_pressure_accel(self._h, _ptr(rho), _ptr(prs), _ptr(acc)))
        return rho, prs, acc

    def neighbours(self, lists=True, source="walk"):
        """source="walk": the validation kernel that re-evaluates the predicate over the 27 cells;
        source="mask": the production hit words of the density pass (what the force pass consumes), decoded."""
        fn = self.lib.sph_download_mask_neighbours if source == "mask" else self.lib.sph_download_neighbours
        n = self.n
        counts = np.empty(n, dtype=np.int32)
        total = C.c_uint64(0)
        self._ck(fn(self._h, _ptr(counts), None, 0, C.byref(total)))
        if not lists:
            return counts, None
        flat = np.empty(total.value, dtype=np.int32)
        self._ck(fn(self._h, _ptr(counts), _ptr(flat), total.value, C.byref(total)))
        return counts, flat

    @property
    def uses_mask_passes(self):
        return self.grid_res[0] >= 4

    def brute_neighbour_counts(self):
        out = np.empty(self.n, dtype=np.int32)
        self._ck(self.lib.sph_brute_neighbour_counts(self._h, _ptr(out)))
        return out

    def stats(self):
        out = np.zeros(6, dtype=np.float64)
        self._ck(self.lib.sph_stats(self._h, _ptr(out)))
        return dict(ke=out[0], com=out[1:4].copy(), fill=out[4], mean_speed=out[5])  # 95th percentile: fill_height_percentile()

    def fill_height_percentile(self, q=0.95):
        out = C.c_double(0)
        self._ck(self.lib.sph_fill_height_percentile(self._h, float(q), C.byref(out)))
        return out.value

    def set_option(self, name, value):
        self._ck(self.lib.sph_set_option(self._h, name.encode(), int(value)))

    def download_owned(self):
        """Owned particles in canonical order (slab mode: this rank's share; records carry the ids)."""
        n = self.n
        out = np.zeros(max(n, 1), dtype=PARTICLE_DTYPE)
        got = C.c_uint32(0)
        self._ck(self.lib.sph_download_owned(self._h, _ptr(out), out.shape[0], C.byref(got)))
        return out[: got.value]

    def slab_info(self):
        out = np.zeros(8, dtype=np.int32)
        rc = self.lib.sph_slab_info(self._h, _ptr(out))
        if rc:
            raise SphError("not a slab context")
        keys = ("rank", "world", "z0", "z1", "z_base", "rz_local", "n_own", "mean_sent_per_step")
        return dict(zip(keys, (int(v) for v in out)))

    def counter(self, name):
        v = C.c_uint64(0)
        rc = self.lib.sph_get_counter(self._h, name.encode(), C.byref(v))
        if rc:
            raise SphError(f"unknown counter {name}")
        return v.value


class _BorrowedContext(SphContext):
    """A context owned by a C++ simulator (never destroyed from Python)."""

    def __init__(self, handle, cfg):
        self.lib = cuda_lib()
        self._h = C.c_void_p(handle)
        self.cfg = cfg

    def close(self):
        self._h = None

    __del__ = close


class _OwnedArray(np.ndarray):
    """ndarray view of memory owned by another Python object (kept alive through `_owner`)."""

    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None and self._owner is None:
            self._owner = getattr(obj, "_owner", None)


class Simulator:
    """The C++ simulator object (CCUDAParticleSimulator / scene-only) through the facade."""

    def __init__(self, kind="cuda", box=0.9, device=0, scenario=DAM_BREAK):
        """kind: "cuda" | "cuda_brute" | "scene_only", or an eSimulationType value (int) for the factory path."""
        self.lib = host_lib()
        if np.isscalar(box):
            box = (box, box, box)
        self.box = tuple(float(np.float32(b)) for b in box)
        self.kind = kind
        if isinstance(kind, int):
            self._h = self.lib.gmu_sim_create_by_type(kind, *self.box, int(device), int(scenario))
        else:
            self._h = self.lib.gmu_sim_create(kind.encode(), *self.box, int(device), int(scenario))
        if not self._h:
            raise SphError(self.lib.gmu_sim_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.gmu_sim_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise SphError(self.lib.gmu_sim_last_error().decode())

    def set_owned_layers(self, z0, z1):
        self._ck(self.lib.gmu_sim_set_owned_layers(self._h, int(z0), int(z1)))
        return self

    def enable_slab(self, rank, world, nccl_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(nccl_id)[:128].ljust(128, b"\0"))
        self._ck(self.lib.gmu_sim_enable_slab(self._h, int(rank), int(world), buf))
        return self

    def setup_scene(self):
        self._ck(self.lib.gmu_sim_setup_scene(self._h))
        return self

    def step(self, n=1):
        self._ck(self.lib.gmu_sim_step(self._h, int(n)))

    def step_many(self, n, timed=True):
        ms = C.c_double(0)
        self._ck(self.lib.gmu_sim_step_many(self._h, int(n), C.byref(ms) if timed else None))
        return ms.value

    def emit(self, n_steps=1):
        self._ck(self.lib.gmu_sim_emit(self._h, int(n_steps)))

    def set_mirror_mode(self, mode):
        self._ck(self.lib.gmu_sim_set_mirror_mode(self._h, int(mode)))

    def set_mirror_stride(self, stride):
        self._ck(self.lib.gmu_sim_set_mirror_stride(self._h, int(stride)))

    def sync_host(self):
        self._ck(self.lib.gmu_sim_sync_host(self._h))

    def wait_host(self):
        self._ck(self.lib.gmu_sim_wait_host(self._h))

    def set_gravity(self, g):
        self._ck(self.lib.gmu_sim_set_gravity(self._h, float(g[0]), float(g[1]), float(g[2])))

    def set_collision_faces(self, faces):
        faces = np.ascontiguousarray(faces, dtype=np.float32).reshape(-1, 12)
        self._ck(self.lib.gmu_sim_set_collision_faces(self._h, _ptr(faces) if faces.shape[0] else None, faces.shape[0]))

    def key(self, qt_key):
        self._ck(self.lib.gmu_sim_key(self._h, int(qt_key)))

    @property
    def gravity(self):
        out = np.zeros(3, dtype=np.float32)
        self._ck(self.lib.gmu_sim_get_gravity(self._h, _ptr(out)))
        return out

    @property
    def running(self):
        return bool(self.lib.gmu_sim_is_running(self._h))

    def push_event(self, iteration, fps, grid, density, forces, collisions, integrate):
        ev = np.array([iteration, fps, grid, density, forces, collisions, integrate], dtype=np.float64)
        self._ck(self.lib.gmu_sim_push_event(self._h, _ptr(ev)))

    def set_profiling(self, on=True, stride=10):
        self._ck(self.lib.gmu_sim_set_profiling(self._h, int(on), int(stride)))

    def set_emission_multiplier(self, nozzles):
        self._ck(self.lib.gmu_sim_set_emission_multiplier(self._h, int(nozzles)))

    @property
    def n(self):
        return int(self.lib.gmu_sim_particle_count(self._h))

    @property
    def max_count(self):
        return int(self.lib.gmu_sim_max_particle_count(self._h))

    @property
    def iteration(self):
        return int(self.lib.gmu_sim_iteration(self._h))

    @property
    def device(self):
        return self.lib.gmu_sim_device_name(self._h).decode()

    def host_particles(self):
        """View of the host mirror m_clParticles (valid until the simulator is destroyed)."""
        n = self.n
        if n == 0:
            return np.zeros(0, dtype=PARTICLE_DTYPE)
        addr = self.lib.gmu_sim_host_particles(self._h)
        buf = (C.c_uint8 * (n * 80)).from_address(addr)
        view = np.frombuffer(buf, dtype=PARTICLE_DTYPE, count=n).view(_OwnedArray)
        view._owner = self  # the memory belongs to the C++ simulator: keep it alive as long as any view of it is
        return view

    def context(self):
        h = self.lib.gmu_sim_context(self._h)
        if not h:
            raise SphError("simulator has no device context (call setup_scene on a CUDA simulator)")
        return _BorrowedContext(h, make_config(self.box, max(self.max_count, 1)))

    def events(self):
        n = int(self.lib.gmu_sim_event_count(self._h))
        out = np.zeros((n, 7), dtype=np.float64)
        if n:
            self.lib.gmu_sim_get_events(self._h, _ptr(out), n)
        return out

    def export_logs(self, directory, name="CUDA Grid"):
        self._ck(self.lib.gmu_sim_export_logs(self._h, directory.encode(), name.encode()))
