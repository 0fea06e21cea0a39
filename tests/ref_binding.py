"""ctypes binding of oracle/_ref/libsph_ref.so — the REFERENCE'S OWN CCPUParticleSimulator, compiled
unmodified from /root/reference against oracle/qt_shim (oracle/Makefile target `ref`, oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_ref_golden.py and bench.py's
--impl reference / cpu_baseline legs.  Never imported by the product package.

The library exists where /root/reference exists (this container) and travels to the GPU box as a built,
git-ignored file; `available()` says whether it can be used, tests skip otherwise and fall back on the
committed golden vectors it produced (tests/golden/ref_*.npz).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "_ref", "libsph_ref.so")
REFERENCE_ROOT = "/root/reference"

DAM_BREAK, FOUNTAIN = 0, 1
KEY_SPACE, KEY_G, KEY_O, KEY_P, KEY_S = 0x20, 0x47, 0x4F, 0x50, 0x53


def build_ref():
    """Compile the reference's sources where they lie (needs /root/reference; never copies them)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), "ref"])
    return _LIB_PATH


def available():
    if os.path.exists(_LIB_PATH):
        return True
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        try:
            build_ref()
        except (subprocess.CalledProcessError, OSError):
            return False
        return os.path.exists(_LIB_PATH)
    return False


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libsph_ref.so is not built and /root/reference is absent")
    L = C.CDLL(_LIB_PATH)
    vp, f, d, i32, i64 = C.c_void_p, C.c_float, C.c_double, C.c_int, C.c_int64
    sig = {
        "ref_create": (vp, [f, i32]),
        "ref_destroy": (None, [vp]),
        "ref_setup_scene": (None, [vp]),
        "ref_step": (None, [vp, i32]),
        "ref_update_grid": (d, [vp]),
        "ref_update_density_pressure": (d, [vp]),
        "ref_update_forces": (d, [vp]),
        "ref_update_collisions": (d, [vp]),
        "ref_integrate": (d, [vp]),
        "ref_set_gravity": (None, [vp, f, f, f]),
        "ref_get_gravity": (None, [vp, vp]),
        "ref_key": (None, [vp, i32]),
        "ref_is_running": (i32, [vp]),
        "ref_last_iteration": (C.c_ulong, []),
        "ref_device_name": (None, [vp, C.c_char_p, i32]),
        "ref_count": (i64, [vp]),
        "ref_max_count": (i64, [vp]),
        "ref_grid_res": (None, [vp, vp]),
        "ref_params": (None, [vp, vp]),
        "ref_dt": (f, [vp]),
        "ref_walls": (None, [vp, vp]),
        "ref_wall_bounce": (None, [vp, vp, vp, vp]),
        "ref_mesh_bounce": (None, [vp, vp, vp, vp]),
        "ref_get_vec": (None, [vp, i32, vp]),
        "ref_get_scalar": (None, [vp, i32, vp]),
        "ref_set_state": (i32, [vp, i64, vp, vp]),
        "ref_get_cells": (None, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Reference:
    """The reference's CCPUParticleSimulator(scene, boxSize, scenario); method names follow oracle_binding.Oracle."""

    def __init__(self, box, scenario=DAM_BREAK):
        self.box = float(np.float32(box))
        self._h = lib().ref_create(self.box, int(scenario))
        self.scenario = scenario

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h)
            self._h = None

    def setup_scene(self):
        lib().ref_setup_scene(self._h)
        return self

    def set_state(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        vel = np.zeros_like(pos) if vel is None else np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
        if lib().ref_set_state(self._h, pos.shape[0], _p(pos), _p(vel)) != 0:
            raise ValueError("set_state: the reference cannot create particles outside its scene generators; "
                             f"n must equal the current count {self.n}")
        return self

    def set_gravity(self, g):
        lib().ref_set_gravity(self._h, *[float(v) for v in g])

    @property
    def gravity(self):
        out = np.zeros(3, dtype=np.float32)
        lib().ref_get_gravity(self._h, _p(out))
        return out

    def key(self, k):
        lib().ref_key(self._h, int(k))

    @property
    def running(self):
        return bool(lib().ref_is_running(self._h))

    @property
    def device_name(self):
        buf = C.create_string_buffer(128)
        lib().ref_device_name(self._h, buf, 128)
        return buf.value.decode()

    def update_grid(self):
        return lib().ref_update_grid(self._h)

    def update_density_pressure(self):
        return lib().ref_update_density_pressure(self._h)

    def update_forces(self):
        return lib().ref_update_forces(self._h)

    def update_collisions(self):
        return lib().ref_update_collisions(self._h)

    def integrate(self):
        return lib().ref_integrate(self._h)

    def step(self, n=1):
        lib().ref_step(self._h, int(n))

    @property
    def n(self):
        return int(lib().ref_count(self._h))

    @property
    def max_count(self):
        return int(lib().ref_max_count(self._h))

    @property
    def grid_res(self):
        r = np.zeros(3, dtype=np.int32)
        lib().ref_grid_res(self._h, _p(r))
        return tuple(int(v) for v in r)

    @property
    def n_cells(self):
        r = self.grid_res
        return r[0] * r[1] * r[2]

    @property
    def params(self):
        """m_systemParams: poly6, spiky, viscosity constants as fp32 (src/CBaseParticleSimulator.cpp:23-25)."""
        out = np.zeros(3, dtype=np.float32)
        lib().ref_params(self._h, _p(out))
        return out

    @property
    def dt(self):
        return np.float32(lib().ref_dt(self._h))

    @property
    def walls(self):
        out = np.zeros((6, 6), dtype=np.float32)
        lib().ref_walls(self._h, _p(out))
        return out

    def wall_bounce(self, pos, vel):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float32)
        out = np.zeros(3, dtype=np.float32)
        lib().ref_wall_bounce(self._h, _p(pos), _p(vel), _p(out))
        return out

    def mesh_bounce(self, pos, vel):
        """CCollisionGeometry::inverseBounce over the cuboid's 12 triangles (prepared, never called by step())."""
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float32)
        out = np.zeros(3, dtype=np.float32)
        lib().ref_mesh_bounce(self._h, _p(pos), _p(vel), _p(out))
        return out

    def _vec(self, what):
        out = np.zeros((self.n, 3), dtype=np.float32)
        lib().ref_get_vec(self._h, what, _p(out))
        return out

    def _scalar(self, what):
        out = np.zeros(self.n, dtype=np.float32)
        lib().ref_get_scalar(self._h, what, _p(out))
        return out

    pos = property(lambda self: self._vec(0))
    vel = property(lambda self: self._vec(1))
    acc = property(lambda self: self._vec(2))
    density = property(lambda self: self._scalar(0))
    pressure = property(lambda self: self._scalar(1))

    def cells(self):
        """(cell_start[cells+1], ids[n]) in the reference's own intra-cell (swap-and-pop history) order."""
        cs = np.empty(self.n_cells + 1, dtype=np.int32)
        ids = np.empty(self.n, dtype=np.int32)
        lib().ref_get_cells(self._h, _p(cs), _p(ids))
        return cs, ids
