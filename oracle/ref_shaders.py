"""ctypes wrapper of oracle/_ref/libref_shaders.so — the reference's OWN GLSL simulation shaders, translated
mechanically from the reference checkout and compiled for the host (oracle/ref_shim/).  TEST INFRASTRUCTURE ONLY:
imported by tests/, tests/golden/make_ref_shader_golden.py and bench.py's CPU legs (cpu_baseline / --impl reference,
where it is the thing timed as "the reference on the host cores"), never by the product or smoke().

The library is BUILT only where the reference checkout is (this container); like the other built .so files it is
git-ignored but travels to the GPU box with the snapshot, where `available()` then finds the prebuilt file.  Without it
the committed golden vectors generated from it stand in (tests/golden/ref_shader_golden.npz) and bench.py times the port.

RefShaderSim has the interface of oracle.OracleSim (same field / pass ids), so a test can drive both side by side."""
from __future__ import annotations

import ctypes
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_ROOT = os.environ.get("WSB_REFERENCE_ROOT", "/root/reference")
_lib = None

FIELD_BASE, FIELD_WATER, FIELD_WALL, FIELD_LIGHT, FIELD_FEEDBACK, FIELD_DEPOSITION, FIELD_CURL, FIELD_VORT = range(8)
_FIELD_DROPS, _FIELD_LIGHTNING = 8, 9
_CHANNELS = {FIELD_BASE: 4, FIELD_WATER: 4, FIELD_LIGHT: 4, FIELD_FEEDBACK: 4, FIELD_DEPOSITION: 2, FIELD_CURL: 1, FIELD_VORT: 2}


def _builder():
    spec = importlib.util.spec_from_file_location("wsb_build_ref", os.path.join(_HERE, "ref_shim", "build_ref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build(force: bool = False):
    """(Re)build from the reference checkout when it is present; returns the library path or None."""
    return _builder().build(_REF_ROOT, force=force)


def available() -> bool:
    try:
        return build() is not None
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libref_shaders.so needs the reference checkout (%s)" % _REF_ROOT)
        L = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.refsim_create.restype = vp
        L.refsim_create.argtypes = [ci, ci, ci]
        L.refsim_destroy.argtypes = [vp]
        L.refsim_upload.argtypes = [vp] * 5
        L.refsim_set_params.argtypes = [vp, vp]
        L.refsim_set_frame_inputs.argtypes = [vp, vp]
        L.refsim_set_profiles.argtypes = [vp] * 5
        L.refsim_step.argtypes = [vp, ci]
        L.refsim_run_pass.argtypes = [vp, ci]
        L.refsim_field_f32.restype = ctypes.POINTER(ctypes.c_float)
        L.refsim_field_f32.argtypes = [vp, ci, ci]
        L.refsim_field_i8.restype = ctypes.POINTER(ctypes.c_int8)
        L.refsim_field_i8.argtypes = [vp, ci]
        L.refsim_get_iter.restype = ctypes.c_long
        L.refsim_get_iter.argtypes = [vp]
        L.refsim_set_iter.argtypes = [vp, ctypes.c_long]
        L.refsim_get_even.argtypes = [vp]
        L.refsim_set_even.argtypes = [vp, ci]
        L.refsim_last_drops.argtypes = [vp]
        L.refsim_set_last_drops.argtypes = [vp, ci]
        L.refsim_get_inactive.restype = ctypes.c_float
        L.refsim_get_inactive.argtypes = [vp]
        L.refsim_set_inactive.argtypes = [vp, ctypes.c_float]
        L.refsim_set_threads.argtypes = [ci]
        L.refsim_setup.argtypes = [ci, ci] + [ctypes.c_float] * 4 + [vp] * 4
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class RefShaderSim:
    """The reference's simulation state + draw loop (app.js:5830-6005) with its own shaders, on the CPU."""

    def __init__(self, width: int, height: int, n_droplets: int = 0):
        self.L = lib()
        self.W, self.H, self.ND = width, height, n_droplets
        h = self.L.refsim_create(width, height, n_droplets)
        if not h:
            raise ValueError("height beyond the per-row uniform tables of the build (WSB_REF_MAX_ROWS, glsl_shim.h)")
        self.h = ctypes.c_void_p(h)

    def close(self):
        if self.h:
            self.L.refsim_destroy(self.h)
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
        self.L.refsim_upload(self.h, _ptr(base), _ptr(water), _ptr(wall), _ptr(drops))

    def set_params(self, p):
        self.L.refsim_set_params(self.h, ctypes.byref(p))

    def set_frame_inputs(self, fi):
        self.L.refsim_set_frame_inputs(self.h, ctypes.byref(fi))

    def set_profiles(self, initial_T, snd_T=None, snd_W=None, snd_Vel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (initial_T, snd_T, snd_W, snd_Vel)]
        for a in arrs:
            assert a is None or a.shape == (self.H + 1,)
        self.L.refsim_set_profiles(self.h, *[_ptr(a) for a in arrs])

    def step(self, n=1):
        self.L.refsim_step(self.h, n)

    def run_pass(self, p):
        self.L.refsim_run_pass(self.h, p)

    def field(self, field, buf=0, copy=True):
        if field == FIELD_WALL:
            a = np.ctypeslib.as_array(self.L.refsim_field_i8(self.h, buf), shape=(self.H, self.W, 4))
        else:
            a = np.ctypeslib.as_array(self.L.refsim_field_f32(self.h, field, buf), shape=(self.H, self.W, _CHANNELS[field]))
        return a.copy() if copy else a

    def droplets(self, buf=None, copy=True):
        if buf is None:
            buf = self.L.refsim_last_drops(self.h)
        a = np.ctypeslib.as_array(self.L.refsim_field_f32(self.h, _FIELD_DROPS, buf), shape=(max(self.ND, 1), 5))[: self.ND]
        return a.copy() if copy else a

    @property
    def lightning(self):
        return np.ctypeslib.as_array(self.L.refsim_field_f32(self.h, _FIELD_LIGHTNING, 0), shape=(4,)).copy()

    @property
    def iter(self):
        return self.L.refsim_get_iter(self.h)

    @iter.setter
    def iter(self, v):
        self.L.refsim_set_iter(self.h, v)

    @property
    def even(self):
        return bool(self.L.refsim_get_even(self.h))

    @property
    def inactive_droplets(self):
        return self.L.refsim_get_inactive(self.h)

    @inactive_droplets.setter
    def inactive_droplets(self, v):
        self.L.refsim_set_inactive(self.h, v)

    def light_latest(self):
        return self.field(FIELD_LIGHT, 0 if self.even else 1)

    def copy_state_from(self, o):
        """Take over the complete state of an oracle.OracleSim (same size): every texture, both droplet buffers,
        the latches and the loop counters — so that one pass can be run on identical inputs in both."""
        assert (o.W, o.H, o.ND) == (self.W, self.H, self.ND)
        for f in (FIELD_BASE, FIELD_WATER, FIELD_WALL, FIELD_LIGHT):
            for b in (0, 1):
                self.field(f, b, copy=False)[...] = o.field(f, b, copy=False)
        for f in (FIELD_FEEDBACK, FIELD_DEPOSITION, FIELD_CURL, FIELD_VORT):
            self.field(f, 0, copy=False)[...] = o.field(f, 0, copy=False)
        if self.ND:
            for b in (0, 1):
                self.droplets(b, copy=False)[...] = o.droplets(b, copy=False)
        np.ctypeslib.as_array(self.L.refsim_field_f32(self.h, _FIELD_LIGHTNING, 0), shape=(4,))[...] = o.lightning
        self.iter = o.iter
        self.inactive_droplets = o.inactive_droplets
        self.L.refsim_set_even(self.h, 1 if o.even else 0)
        self.L.refsim_set_last_drops(self.h, o.L.oracle_last_drops(o.h))


def setup_state(width: int, height: int, seed: float, height_mult: float, sim_height: float, dry_lapse: float, initial_T):
    """setupShader.frag drawn once into frameBuff_0 (app.js:5729-5742): (base, water, wall)."""
    initial_T = np.ascontiguousarray(initial_T, np.float32)
    assert initial_T.shape == (height + 1,)
    base = np.zeros((height, width, 4), np.float32)
    water = np.zeros((height, width, 4), np.float32)
    wall = np.zeros((height, width, 4), np.int8)
    lib().refsim_setup(width, height, seed, height_mult, sim_height, dry_lapse, _ptr(initial_T), _ptr(base), _ptr(water), _ptr(wall))
    return base, water, wall
