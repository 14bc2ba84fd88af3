"""ctypes wrapper of the CPU oracle (oracle/wsb_oracle.cpp).  TEST INFRASTRUCTURE ONLY — imported
by tests/, bench.py's cpu_baseline / `--impl reference` leg and __graft_entry__.smoke(); never by
the product package."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

FIELD_BASE, FIELD_WATER, FIELD_WALL, FIELD_LIGHT, FIELD_FEEDBACK, FIELD_DEPOSITION, FIELD_CURL, FIELD_VORT = range(8)
_FIELD_DROPS, _FIELD_LIGHTNING = 8, 9
PASS_VELOCITY, PASS_CURL, PASS_VORTICITY, PASS_BOUNDARY, PASS_ADVECTION, PASS_PRESSURE, PASS_LIGHTING, PASS_PRECIPITATION, PASS_ITER_INC, PASS_ADVECTION_DRY = range(10)
_CHANNELS = {FIELD_BASE: 4, FIELD_WATER: 4, FIELD_LIGHT: 4, FIELD_FEEDBACK: 4, FIELD_DEPOSITION: 2, FIELD_CURL: 1, FIELD_VORT: 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "wsb_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.c_int] * 5
        L.oracle_field_f32.restype = ctypes.POINTER(ctypes.c_float)
        L.oracle_field_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.oracle_field_i8.restype = ctypes.POINTER(ctypes.c_int8)
        L.oracle_field_i8.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for name in ("oracle_destroy",):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.oracle_upload.argtypes = [ctypes.c_void_p] * 5
        L.oracle_set_params.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_set_frame_inputs.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_set_profiles.argtypes = [ctypes.c_void_p] * 5
        L.oracle_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_step_dry.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_run_pass.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_get_iter.restype = ctypes.c_long
        L.oracle_get_iter.argtypes = [ctypes.c_void_p]
        L.oracle_set_iter.argtypes = [ctypes.c_void_p, ctypes.c_long]
        L.oracle_get_even.argtypes = [ctypes.c_void_p]
        L.oracle_last_drops.argtypes = [ctypes.c_void_p]
        L.oracle_get_inactive.restype = ctypes.c_float
        L.oracle_get_inactive.argtypes = [ctypes.c_void_p]
        L.oracle_set_inactive.argtypes = [ctypes.c_void_p, ctypes.c_float]
        for name in ("oracle_maxWater", "oracle_IR_emitted", "oracle_cbrt"):
            getattr(L, name).restype = ctypes.c_float
            getattr(L, name).argtypes = [ctypes.c_float]
        L.oracle_random2d.restype = ctypes.c_float
        L.oracle_random2d.argtypes = [ctypes.c_float, ctypes.c_float]
        L.oracle_map_rangeC.restype = ctypes.c_float
        L.oracle_map_rangeC.argtypes = [ctypes.c_float] * 5
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_get_threads.restype = ctypes.c_int
        L.oracle_hash.restype = ctypes.c_uint32
        L.oracle_hash.argtypes = [ctypes.c_uint32]
        _lib = L
    return _lib


def set_threads(n: int | None = None) -> int:
    """Use n OpenMP threads (default: every core this process may run on); returns the team size."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_threads(int(n))
    return lib().oracle_get_threads()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleSim:
    """Mirror of the reference's simulation state + loop on the CPU."""

    def __init__(self, width: int, height: int, n_droplets: int = 0, global_width: int = 0, x0: int = 0):
        self.L = lib()
        self.W, self.H, self.ND = width, height, n_droplets
        self.h = ctypes.c_void_p(self.L.oracle_create(width, height, n_droplets, global_width, x0))

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, base, water, wall, drops=None):
        base = np.ascontiguousarray(base, np.float32)
        water = np.ascontiguousarray(water, np.float32)
        wall = np.ascontiguousarray(wall, np.int8)
        assert base.shape == (self.H, self.W, 4) and water.shape == base.shape and wall.shape == base.shape
        if drops is not None:
            drops = np.ascontiguousarray(drops, np.float32)
            assert drops.shape == (self.ND, 5)
        self.L.oracle_upload(self.h, _ptr(base), _ptr(water), _ptr(wall), _ptr(drops))

    def set_params(self, p):
        self.L.oracle_set_params(self.h, ctypes.byref(p))

    def set_frame_inputs(self, fi):
        self.L.oracle_set_frame_inputs(self.h, ctypes.byref(fi))

    def set_profiles(self, initial_T, snd_T=None, snd_W=None, snd_Vel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (initial_T, snd_T, snd_W, snd_Vel)]
        for a in arrs:
            assert a is None or a.shape == (self.H + 1,)
        self.L.oracle_set_profiles(self.h, *[_ptr(a) for a in arrs])

    def step(self, n=1):
        self.L.oracle_step(self.h, n)

    def step_dry(self, n=1):
        self.L.oracle_step_dry(self.h, n)

    def run_pass(self, p):
        self.L.oracle_run_pass(self.h, p)

    def field(self, field, buf=0, copy=True):
        n = self.W * self.H
        if field == FIELD_WALL:
            p = self.L.oracle_field_i8(self.h, buf)
            a = np.ctypeslib.as_array(p, shape=(self.H, self.W, 4))
        else:
            ch = _CHANNELS[field]
            p = self.L.oracle_field_f32(self.h, field, buf)
            a = np.ctypeslib.as_array(p, shape=(self.H, self.W, ch))
        return a.copy() if copy else a

    def droplets(self, buf=None, copy=True):
        if buf is None:
            buf = self.L.oracle_last_drops(self.h)
        p = self.L.oracle_field_f32(self.h, _FIELD_DROPS, buf)
        a = np.ctypeslib.as_array(p, shape=(self.ND, 5))
        return a.copy() if copy else a

    @property
    def lightning(self):
        p = self.L.oracle_field_f32(self.h, _FIELD_LIGHTNING, 0)
        return np.ctypeslib.as_array(p, shape=(4,)).copy()

    @property
    def iter(self):
        return self.L.oracle_get_iter(self.h)

    @iter.setter
    def iter(self, v):
        self.L.oracle_set_iter(self.h, v)

    @property
    def even(self):
        return bool(self.L.oracle_get_even(self.h))

    @property
    def inactive_droplets(self):
        return self.L.oracle_get_inactive(self.h)

    @inactive_droplets.setter
    def inactive_droplets(self, v):
        self.L.oracle_set_inactive(self.h, v)

    def light_latest(self):
        """The light texture written by the last lighting pass (even toggles after writing)."""
        return self.field(FIELD_LIGHT, 0 if self.even else 1)
