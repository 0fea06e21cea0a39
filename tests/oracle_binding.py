"""ctypes binding of the CPU oracle (oracle/_build/libsph_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "_build", "libsph_oracle.so")

DAM_BREAK, FOUNTAIN = 0, 1


def build_oracle():
    """Compile the oracle with its own Makefile (g++ -O2 -ffp-contract=off)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return _LIB_PATH


def _load():
    if not os.path.exists(_LIB_PATH):
        build_oracle()
    lib = C.CDLL(_LIB_PATH)
    vp, f, d, i32, i64 = C.c_void_p, C.c_float, C.c_double, C.c_int, C.c_int64
    sig = {
        "oracle_create": (vp, [f, f, f, i32]),
        "oracle_destroy": (None, [vp]),
        "oracle_setup_scene": (None, [vp]),
        "oracle_set_state": (None, [vp, i64, vp, vp]),
        "oracle_overwrite_state": (i32, [vp, i64, vp, vp]),
        "oracle_add_particle": (None, [vp, f, f, f, f, f, f]),
        "oracle_set_gravity": (None, [vp, f, f, f]),
        "oracle_set_faces": (None, [vp, i32, vp]),
        "oracle_set_threads": (i32, [vp, i32]),
        "oracle_generate_particles": (i32, [vp]),
        "oracle_update_grid": (d, [vp]),
        "oracle_update_density_pressure": (d, [vp]),
        "oracle_update_forces": (d, [vp]),
        "oracle_update_collisions": (d, [vp]),
        "oracle_integrate": (d, [vp]),
        "oracle_step": (None, [vp, i32, vp]),
        "oracle_brute_density_pressure": (None, [vp]),
        "oracle_brute_forces": (None, [vp]),
        "oracle_count": (i64, [vp]),
        "oracle_max_count": (i64, [vp]),
        "oracle_grid_res": (None, [vp, vp]),
        "oracle_constants": (None, [vp, vp, vp, vp, vp]),
        "oracle_get_pos": (None, [vp, vp]),
        "oracle_get_vel": (None, [vp, vp]),
        "oracle_get_acc": (None, [vp, vp]),
        "oracle_get_acc_sph": (None, [vp, vp]),
        "oracle_get_acc_wall": (None, [vp, vp]),
        "oracle_get_acc_mesh": (None, [vp, vp]),
        "oracle_get_acc_scale": (None, [vp, vp]),
        "oracle_get_density": (None, [vp, vp]),
        "oracle_get_pressure": (None, [vp, vp]),
        "oracle_get_keys": (None, [vp, vp]),
        "oracle_get_cells": (None, [vp, vp, vp]),
        "oracle_get_cells_raw": (None, [vp, vp, vp]),
        "oracle_get_neighbours": (i64, [vp, vp, vp]),
        "oracle_stats": (None, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Thin object wrapper; method names follow the reference's phase names."""

    def __init__(self, box, scenario=DAM_BREAK):
        if np.isscalar(box):
            box = (box, box, box)
        self.box = tuple(float(np.float32(b)) for b in box)
        self._h = lib().oracle_create(*self.box, int(scenario))
        self.scenario = scenario

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    # scene
    def setup_scene(self):
        lib().oracle_setup_scene(self._h)
        return self

    def set_state(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        vel = np.zeros_like(pos) if vel is None else np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
        lib().oracle_set_state(self._h, pos.shape[0], _p(pos), _p(vel))
        return self

    def overwrite_state(self, pos, vel):
        """New positions / velocities for the existing particles; cell vectors keep their history."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        vel = np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
        if lib().oracle_overwrite_state(self._h, pos.shape[0], _p(pos), _p(vel)) != 0:
            raise ValueError("overwrite_state: n must equal the current count")
        return self

    def add_particle(self, x, y, z, vx=0.0, vy=0.0, vz=0.0):
        lib().oracle_add_particle(self._h, x, y, z, vx, vy, vz)

    def set_gravity(self, g):
        lib().oracle_set_gravity(self._h, *[float(v) for v in g])

    def set_faces(self, faces):
        """(n, 12) float32: normal, v0, v1, v2 per face; empty clears the mesh (shipped behaviour)."""
        faces = np.ascontiguousarray(faces, dtype=np.float32).reshape(-1, 12)
        lib().oracle_set_faces(self._h, faces.shape[0], _p(faces) if faces.shape[0] else None)

    def set_threads(self, threads=0):
        """NOT reference behaviour: worker threads for the labelled all-cores baseline (0 = all cores)."""
        return lib().oracle_set_threads(self._h, int(threads))

    def generate_particles(self):
        return lib().oracle_generate_particles(self._h)

    # phases
    def update_grid(self):
        return lib().oracle_update_grid(self._h)

    def update_density_pressure(self):
        return lib().oracle_update_density_pressure(self._h)

    def update_forces(self):
        return lib().oracle_update_forces(self._h)

    def update_collisions(self):
        return lib().oracle_update_collisions(self._h)

    def integrate(self):
        return lib().oracle_integrate(self._h)

    def step(self, n=1):
        ms = np.zeros(5, dtype=np.float64)
        lib().oracle_step(self._h, int(n), _p(ms))
        return ms

    def brute_density_pressure(self):
        lib().oracle_brute_density_pressure(self._h)

    def brute_forces(self):
        lib().oracle_brute_forces(self._h)

    # taps
    @property
    def n(self):
        return int(lib().oracle_count(self._h))

    @property
    def max_count(self):
        return int(lib().oracle_max_count(self._h))

    @property
    def grid_res(self):
        r = np.zeros(3, dtype=np.int32)
        lib().oracle_grid_res(self._h, _p(r))
        return tuple(int(v) for v in r)

    @property
    def n_cells(self):
        r = self.grid_res
        return r[0] * r[1] * r[2]

    def constants(self):
        a, b, c, h2 = C.c_double(), C.c_double(), C.c_double(), C.c_float()
        lib().oracle_constants(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(h2))
        return a.value, b.value, c.value, np.float32(h2.value)

    def _vec(self, fn, width):
        out = np.empty((self.n, width) if width > 1 else (self.n,), dtype=np.float32)
        getattr(lib(), fn)(self._h, _p(out))
        return out

    pos = property(lambda self: self._vec("oracle_get_pos", 3))
    vel = property(lambda self: self._vec("oracle_get_vel", 3))
    acc = property(lambda self: self._vec("oracle_get_acc", 3))
    acc_sph = property(lambda self: self._vec("oracle_get_acc_sph", 3))
    acc_wall = property(lambda self: self._vec("oracle_get_acc_wall", 3))
    acc_mesh = property(lambda self: self._vec("oracle_get_acc_mesh", 3))
    acc_scale = property(lambda self: self._vec("oracle_get_acc_scale", 1))
    density = property(lambda self: self._vec("oracle_get_density", 1))
    pressure = property(lambda self: self._vec("oracle_get_pressure", 1))

    def keys(self):
        out = np.empty(self.n, dtype=np.int32)
        lib().oracle_get_keys(self._h, _p(out))
        return out

    def cells(self):
        """(cell_start[cells+1], ids[n]) — canonical (cell, id) order."""
        cs = np.empty(self.n_cells + 1, dtype=np.int32)
        ids = np.empty(self.n, dtype=np.int32)
        lib().oracle_get_cells(self._h, _p(cs), _p(ids))
        return cs, ids

    def cells_raw(self):
        """(cell_start[cells+1], ids[n]) in the cell vectors' own (swap-and-pop history) order."""
        cs = np.empty(self.n_cells + 1, dtype=np.int32)
        ids = np.empty(self.n, dtype=np.int32)
        lib().oracle_get_cells_raw(self._h, _p(cs), _p(ids))
        return cs, ids

    def neighbours(self, lists=True):
        counts = np.empty(self.n, dtype=np.int32)
        total = lib().oracle_get_neighbours(self._h, _p(counts), None)
        if not lists:
            return counts, None
        flat = np.empty(int(total), dtype=np.int32)
        lib().oracle_get_neighbours(self._h, _p(counts), _p(flat))
        return counts, flat

    def stats(self):
        out = np.zeros(6, dtype=np.float64)
        lib().oracle_stats(self._h, _p(out))
        return dict(ke=out[0], com=out[1:4].copy(), fill=out[4], fill95=out[5])
