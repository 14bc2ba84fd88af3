"""Host-side mirror of the reference's simulation surface, bound to libwsb200.so over the C ABI.

The reference has no operator/plugin API for this path: `draw()` (app.js:5686) owns the WebGL
objects and runs the loop inline.  `Simulation` reproduces the interactions `app.js` has with the
simulation state, under the reference's own names where it has them:

    reference (app.js)                                   here
    -------------------------------------------------    ------------------------------------
    loadData() + mainScript(base, water, wall, drops)    Simulation.from_save / Simulation(...)
    setGuiUniforms()                     3401-3443       set_gui_uniforms(guiControls)
    updateSunlight()                     6510-6572       update_sunlight(...)
    for (i < IterPerFrame) {...}         5830-6005       step(IterPerFrame)
    gl.readPixels(x, y, w, h, ...)       25 call sites   read_pixels(field, x, y, w, h)
    getBufferSubData(precipVertexBuffer) 5017, 6595      read_droplets()
    prepareDownload()                    6575-6628       prepare_download() -> SaveFile

There is no CPU fallback: every method calls into the CUDA library and raises if it is missing.
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

from . import params as P
from . import savefile as S
from . import strips

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELD_BASE, FIELD_WATER, FIELD_WALL, FIELD_LIGHT, FIELD_FEEDBACK, FIELD_DEPOSITION, FIELD_CURL, FIELD_VORTFORCE = range(8)
VIEW_FRAMEBUFF_0, VIEW_FRAMEBUFF_1, VIEW_LATEST = range(3)
SCHEDULE_FUSED, SCHEDULE_REFERENCE = 0, 1
(PASS_VELOCITY, PASS_CURL, PASS_VORTICITY, PASS_BOUNDARY, PASS_ADVECTION, PASS_PRESSURE, PASS_LIGHTING,
 PASS_PRECIPITATION, PASS_ITER_INC, PASS_ADVECTION_DRY) = range(10)
KERNEL_PVB, KERNEL_ADV, KERNEL_DRY, KERNEL_PRECIP, KERNEL_HALO, KERNEL_WAIT, KERNEL_EDGE, KERNEL_SPRITES = range(8)
COMM_ID_BYTES = 128
PEER_INFO_BYTES = 256
ABI_VERSION = 1

_FIELD_SPEC = {  # channels, dtype
    FIELD_BASE: (4, np.float32), FIELD_WATER: (4, np.float32), FIELD_WALL: (4, np.int8),
    FIELD_LIGHT: (4, np.float32), FIELD_FEEDBACK: (4, np.float32), FIELD_DEPOSITION: (2, np.float32),
    FIELD_CURL: (1, np.float32), FIELD_VORTFORCE: (2, np.float32),
}


class WsbConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("n_droplets", ctypes.c_int32), ("device", ctypes.c_int32), ("rank", ctypes.c_int32),
        ("n_ranks", ctypes.c_int32), ("schedule", ctypes.c_int32),
        ("comm_id", ctypes.c_uint8 * COMM_ID_BYTES),
    ]


class WsbError(RuntimeError):
    pass


def library_path() -> str:
    """csrc/libwsb200.so; WSB200_LIB overrides it (A/B timing of experimental builds)."""
    return os.environ.get("WSB200_LIB") or os.path.join(_HERE, "csrc", "libwsb200.so")


# every symbol include/wsb200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "wsb_comm_id_create", "wsb_peer_info", "wsb_connect_peers", "wsb_set_exchange", "wsb_create", "wsb_destroy", "wsb_upload", "wsb_upload_local", "wsb_get_layout",
    "wsb_set_profiling", "wsb_kernel_time_ms", "wsb_set_params",
    "wsb_set_profiles", "wsb_set_frame_inputs", "wsb_step", "wsb_sync", "wsb_debug_run_pass",
    "wsb_step_dry", "wsb_read_rect", "wsb_read_points", "wsb_read_droplets", "wsb_get_inactive_droplets",
    "wsb_get_lightning", "wsb_get_iter", "wsb_set_iter", "wsb_get_strip", "wsb_get_max_velocity",
    "wsb_get_launch_count", "wsb_count_nonfinite", "wsb_last_step_ms", "wsb_last_error", "wsb_build_info",
]


def load_library():
    """Load libwsb200.so.  Fails loudly when the CUDA extension has not been built — there is no
    other implementation behind this package."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise WsbError(
            f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C 2d-weather-sandbox_b200/csrc). wsb200 has no CPU fallback.")
    L = ctypes.CDLL(path)
    vp, i32, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_float)
    L.wsb_comm_id_create.argtypes = [ctypes.POINTER(ctypes.c_uint8)]
    L.wsb_create.argtypes = [ctypes.POINTER(WsbConfig), ctypes.POINTER(vp)]
    L.wsb_peer_info.argtypes = [vp, ctypes.POINTER(ctypes.c_uint8)]
    L.wsb_connect_peers.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
    L.wsb_set_exchange.argtypes = [vp, i32]
    L.wsb_destroy.argtypes = [vp]
    L.wsb_upload.argtypes = [vp, vp, vp, vp, vp]
    L.wsb_upload_local.argtypes = [vp, vp, vp, vp, vp]
    L.wsb_get_layout.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.wsb_set_profiling.argtypes = [vp, i32]
    L.wsb_kernel_time_ms.argtypes = [vp, i32, f32p, ctypes.POINTER(i32)]
    L.wsb_set_params.argtypes = [vp, ctypes.POINTER(P.WsbParams)]
    L.wsb_set_profiles.argtypes = [vp, vp, vp, vp, vp]
    L.wsb_set_frame_inputs.argtypes = [vp, ctypes.POINTER(P.WsbFrameInputs)]
    L.wsb_step.argtypes = [vp, i32]
    L.wsb_sync.argtypes = [vp]
    L.wsb_debug_run_pass.argtypes = [vp, i32]
    L.wsb_step_dry.argtypes = [vp, i32]
    L.wsb_read_rect.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp]
    L.wsb_read_points.argtypes = [vp, i32, i32, i32, vp, vp]
    L.wsb_read_droplets.argtypes = [vp, i32, i32, i32, vp]
    L.wsb_get_inactive_droplets.argtypes = [vp, f32p]
    L.wsb_get_lightning.argtypes = [vp, f32p]
    L.wsb_get_iter.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
    L.wsb_set_iter.argtypes = [vp, ctypes.c_int64]
    L.wsb_get_strip.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.wsb_get_max_velocity.argtypes = [vp, f32p]
    L.wsb_get_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
    L.wsb_count_nonfinite.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
    L.wsb_last_step_ms.argtypes = [vp, f32p]
    L.wsb_last_error.restype = ctypes.c_char_p
    L.wsb_build_info.restype = ctypes.c_char_p
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def comm_id_create() -> bytes:
    L = load_library()
    buf = (ctypes.c_uint8 * COMM_ID_BYTES)()
    if L.wsb_comm_id_create(buf) != 0:
        raise WsbError(L.wsb_last_error().decode())
    return bytes(buf)


class Simulation:
    """One simulation (one process / GPU; `n_ranks` processes form an x-strip ring)."""

    def __init__(self, width: int, height: int, n_droplets: int | None = None, *, device: int = 0,
                 rank: int = 0, n_ranks: int = 1, comm_id: bytes | None = None,
                 schedule: int = SCHEDULE_FUSED, gui_controls: dict | None = None):
        self.L = load_library()
        self.W, self.H = int(width), int(height)
        self.ND = S.num_droplets(width, height) if n_droplets is None else int(n_droplets)
        self.rank, self.n_ranks = rank, n_ranks
        cfg = WsbConfig()
        cfg.abi_version = ABI_VERSION
        cfg.width, cfg.height, cfg.n_droplets = self.W, self.H, self.ND
        cfg.device, cfg.rank, cfg.n_ranks, cfg.schedule = device, rank, n_ranks, schedule
        if n_ranks > 1 and comm_id is not None:  # None: no NCCL communicator, the strip is linked with connect_peers()
            if len(comm_id) != COMM_ID_BYTES:
                raise WsbError("comm_id must be the 128 bytes created on rank 0")
            ctypes.memmove(cfg.comm_id, comm_id, COMM_ID_BYTES)
        self.h = ctypes.c_void_p()
        self._check(self.L.wsb_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        self.x_begin, self.local_width = strips.strip_bounds(self.W, n_ranks, rank)
        self.gui = dict(gui_controls) if gui_controls is not None else P.resolve_settings(None)
        self.sun_clock = None
        self._soundings = (None, None, None)  # realWorldSounding_{T,W,Vel}v as last set (set_profiles)
        self.set_gui_uniforms(self.gui)
        self.update_sunlight()

    # -- construction helpers ------------------------------------------------------------------
    @classmethod
    def from_save(cls, sf: S.SaveFile, **kw) -> "Simulation":
        """loadData -> mainScript (app.js:1256-1366, 1495): settings resolved with the reference's
        -1 -> default rule, both ping-pong copies initialised from the file."""
        g = P.resolve_settings(sf.settings_json)
        sim = cls(sf.width, sf.height, sf.droplets.shape[0], gui_controls=g, **kw)
        sim.upload(sf.base, sf.water, sf.wall, sf.droplets)
        sim._stations = sf.stations
        return sim

    @classmethod
    def new_simulation(cls, width: int, height: int, seed: float = 0.37, height_mult: float = 0.5,
                       sim_height: float = 12000.0, **kw) -> "Simulation":
        """'Create new simulation' (app.js:1357-1364 + setupShader.frag): default settings with the
        chosen simulation height, terrain from `seed` / `height_mult`, all droplets inactive."""
        from . import synth

        g = P.resolve_settings(None, sim_height=sim_height)
        base, water, wall, drops = synth.setup_state(width, height, seed, height_mult, g)
        sim = cls(width, height, drops.shape[0], gui_controls=g, **kw)
        sim.upload(base, water, wall, drops)
        return sim

    def _check(self, rc: int):
        if rc != 0:
            raise WsbError(self.L.wsb_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.wsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-GPU ghost exchange over peer memory ---------------------------------------------
    def peer_info(self) -> bytes:
        """This strip's exchange window (IPC handle + layout) for its two ring neighbours."""
        buf = (ctypes.c_uint8 * PEER_INFO_BYTES)()
        self._check(self.L.wsb_peer_info(self.h, buf))
        return bytes(buf)

    def connect_peers(self, left_info: bytes, right_info: bytes):
        """Map the neighbours' windows: from now on ghost columns travel as direct peer stores."""
        if len(left_info) != PEER_INFO_BYTES or len(right_info) != PEER_INFO_BYTES:
            raise WsbError("connect_peers: blobs must come from peer_info()")
        self._check(self.L.wsb_connect_peers(self.h, left_info, right_info))

    def set_exchange(self, transport: str):
        """Select the ghost-exchange transport of a strip: "peer" (stores straight into the neighbours' ghost columns),
        "peerc" (the same through a compact landing zone + a local copy) or "nccl"."""
        self._check(self.L.wsb_set_exchange(self.h, {"nccl": 0, "peer": 1, "peerc": 2}[transport]))
        self.transport = transport

    # -- state ---------------------------------------------------------------------------------
    def upload(self, base, water, wall, drops=None):
        base = np.ascontiguousarray(base, np.float32)
        water = np.ascontiguousarray(water, np.float32)
        wall = np.ascontiguousarray(wall, np.int8)
        if base.shape != (self.H, self.W, 4) or water.shape != base.shape or wall.shape != base.shape:
            raise WsbError(f"upload: arrays must be [{self.H}][{self.W}][4]")
        if self.ND:
            if drops is None:
                raise WsbError("upload: droplets required")
            drops = np.ascontiguousarray(drops, np.float32)
            if drops.shape != (self.ND, 5):
                raise WsbError(f"upload: droplets must be [{self.ND}][5]")
        else:
            drops = None
        self._check(self.L.wsb_upload(self.h, _ptr(base), _ptr(water), _ptr(wall), _ptr(drops)))

    def upload_local(self, base, water, wall, drops=None):
        """Upload only this rank's padded strip: arrays [H][ghost + local_width + ghost][4] whose
        column i is global column (x_begin - ghost + i) mod W (see `padded_columns`)."""
        base = np.ascontiguousarray(base, np.float32)
        water = np.ascontiguousarray(water, np.float32)
        wall = np.ascontiguousarray(wall, np.int8)
        x0, lw, gh = self.layout()
        shape = (self.H, lw + 2 * gh, 4)
        if base.shape != shape or water.shape != shape or wall.shape != shape:
            raise WsbError(f"upload_local: arrays must be {shape}")
        if self.ND:
            if drops is None:
                raise WsbError("upload_local: droplets required")
            drops = np.ascontiguousarray(drops, np.float32)
        else:
            drops = None
        self._check(self.L.wsb_upload_local(self.h, _ptr(base), _ptr(water), _ptr(wall), _ptr(drops)))

    def layout(self) -> tuple[int, int, int]:
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._check(self.L.wsb_get_layout(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def padded_columns(self) -> np.ndarray:
        x0, lw, gh = self.layout()
        return (np.arange(x0 - gh, x0 + lw + gh) % self.W).astype(np.int64)

    def set_profiling(self, on: bool = True):
        self._check(self.L.wsb_set_profiling(self.h, 1 if on else 0))

    def kernel_time_ms(self, kernel: int) -> tuple[float, int]:
        """(summed device ms, launches) of one kernel class in the most recent step call."""
        ms, n = ctypes.c_float(), ctypes.c_int32()
        self._check(self.L.wsb_kernel_time_ms(self.h, int(kernel), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def set_gui_uniforms(self, g: dict | None = None):
        """setGuiUniforms + static uniforms + profile arrays (app.js:3401-3443, 5439-5502)."""
        if g is not None:
            self.gui = g
        p = P.derive_params(self.gui)
        self.params = p
        self._check(self.L.wsb_set_params(self.h, ctypes.byref(p)))
        # setGuiUniforms does not touch the realWorldSounding_* uniforms: pass the last sounding again
        t0 = P.initial_T_profile(self.H, self.gui)
        self._check(self.L.wsb_set_profiles(self.h, _ptr(t0), *[_ptr(a) for a in self._soundings]))

    def set_params(self, p: P.WsbParams):
        self.params = p
        self._check(self.L.wsb_set_params(self.h, ctypes.byref(p)))

    def set_profiles(self, initial_T, snd_T=None, snd_W=None, snd_Vel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (initial_T, snd_T, snd_W, snd_Vel)]
        for a in arrs:
            if a is not None and a.shape != (self.H + 1,):
                raise WsbError("profiles must have height+1 entries")
        self._soundings = tuple(arrs[1:])
        self._check(self.L.wsb_set_profiles(self.h, *[_ptr(a) for a in arrs]))

    def update_sunlight(self, delta_hours: float | None = None):
        """updateSunlight (app.js:6510-6572): with a day/night cycle advance the clock by
        delta_hours (the reference calls this once per frame with timePerIteration*IterPerFrame);
        otherwise 'MANUAL_ANGLE' from guiControls.sunAngle."""
        if self.gui.get("dayNightCycle"):
            if self.sun_clock is None:
                self.sun_clock = P.SunClock(self.gui)
            if delta_hours:
                self.sun_clock.advance(delta_hours)
        fresh = P.frame_inputs(self.gui)
        cur = getattr(self, "frame_inputs", None)
        if cur is None:
            self.frame_inputs = fresh
        else:  # updateSunlight only sets sunAngle / sunIntensity (app.js:6557-6561): brush and airplane inputs stay
            cur.sunAngle, cur.sunIntensity = fresh.sunAngle, fresh.sunIntensity
        self._check(self.L.wsb_set_frame_inputs(self.h, ctypes.byref(self.frame_inputs)))

    def set_frame_inputs(self, fi: P.WsbFrameInputs):
        self.frame_inputs = fi
        self._check(self.L.wsb_set_frame_inputs(self.h, ctypes.byref(fi)))

    # -- the hot path --------------------------------------------------------------------------
    def step(self, n_iters: int = 1):
        """The simulation loop body of draw(), n_iters (= IterPerFrame) times. Asynchronous."""
        self._check(self.L.wsb_step(self.h, int(n_iters)))

    def step_dry(self, n_iters: int = 1):
        self._check(self.L.wsb_step_dry(self.h, int(n_iters)))

    def frame(self, iter_per_frame: int | None = None):
        """One draw() worth of simulation (app.js:5814-6005): sun update once per frame, then
        IterPerFrame iterations."""
        n = int(iter_per_frame if iter_per_frame is not None else self.gui["IterPerFrame"])
        if self.gui.get("dayNightCycle"):
            self.update_sunlight(P.TIME_PER_ITERATION * n)
        self.step(n)

    def run_pass(self, p: int):
        self._check(self.L.wsb_debug_run_pass(self.h, int(p)))

    def sync(self):
        self._check(self.L.wsb_sync(self.h))

    # -- readbacks -----------------------------------------------------------------------------
    def read_pixels(self, field: int, x: int = 0, y: int = 0, w: int | None = None, h: int | None = None,
                    view: int = VIEW_FRAMEBUFF_0, out: np.ndarray | None = None) -> np.ndarray:
        """gl.readPixels on one of the simulation textures.  In a multi-GPU run only the columns of
        this rank's strip are filled (the rest of `out` is left as passed in / zero)."""
        w = self.W - x if w is None else w
        h = self.H - y if h is None else h
        ch, dt = _FIELD_SPEC[field]
        if out is None:
            out = np.zeros((h, w, ch), dt)
        elif out.shape != (h, w, ch) or out.dtype != dt or not out.flags.c_contiguous:
            raise WsbError("read_pixels: bad output array")
        self._check(self.L.wsb_read_rect(self.h, field, view, x, y, w, h, _ptr(out)))
        return out

    def read_points(self, field: int, points, view: int = VIEW_FRAMEBUFF_0) -> np.ndarray:
        """Single texels of BASE / WATER / LIGHT at many (x, y) cells in one call (weather stations,
        app.js:1082-1176).  Returns float32 [n][4]."""
        pts = np.ascontiguousarray(points, np.int32).reshape(-1, 2)
        out = np.zeros((pts.shape[0], 4), np.float32)
        if pts.shape[0]:
            self._check(self.L.wsb_read_points(self.h, field, view, pts.shape[0], _ptr(pts), _ptr(out)))
        return out

    def read_droplets(self, buffer: int = 2, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.ND - first if count is None else count
        out = np.zeros((count, 5), np.float32)
        if count:
            self._check(self.L.wsb_read_droplets(self.h, buffer, first, count, _ptr(out)))
        return out

    @property
    def inactive_droplets(self) -> float:
        v = ctypes.c_float()
        self._check(self.L.wsb_get_inactive_droplets(self.h, ctypes.byref(v)))
        return v.value

    @property
    def lightning(self) -> np.ndarray:
        v = (ctypes.c_float * 4)()
        self._check(self.L.wsb_get_lightning(self.h, v))
        return np.array(v[:], np.float32)

    @property
    def iter_num(self) -> int:
        v = ctypes.c_int64()
        self._check(self.L.wsb_get_iter(self.h, ctypes.byref(v)))
        return v.value

    @iter_num.setter
    def iter_num(self, v: int):
        self._check(self.L.wsb_set_iter(self.h, int(v)))

    @property
    def max_velocity(self) -> float:
        v = ctypes.c_float()
        self._check(self.L.wsb_get_max_velocity(self.h, ctypes.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        v = ctypes.c_int64()
        self._check(self.L.wsb_get_launch_count(self.h, ctypes.byref(v)))
        return v.value

    def count_nonfinite(self) -> int:
        """NaN / Inf values in the current base and water fields (debug aid)."""
        v = ctypes.c_int64()
        self._check(self.L.wsb_count_nonfinite(self.h, ctypes.byref(v)))
        return v.value

    def last_step_ms(self) -> float:
        v = ctypes.c_float()
        self._check(self.L.wsb_last_step_ms(self.h, ctypes.byref(v)))
        return v.value

    def strip(self) -> tuple[int, int]:
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self._check(self.L.wsb_get_strip(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # -- save ----------------------------------------------------------------------------------
    def prepare_download(self) -> S.SaveFile:
        """prepareDownload (app.js:6575-6628): frameBuff_0 (base after pressure, water after the
        boundary pass, wall_0) + droplet buffer 0 + stations + guiControls JSON."""
        if self.n_ranks != 1:
            raise WsbError("prepare_download on a strip: gather with read_pixels per rank instead")
        base = self.read_pixels(FIELD_BASE)
        water = self.read_pixels(FIELD_WATER)
        wall = self.read_pixels(FIELD_WALL)
        drops = self.read_droplets(buffer=0)
        st = getattr(self, "_stations", np.zeros((0, 2), np.int16))
        return S.SaveFile(self.W, self.H, base, water, wall, drops, st, json.dumps(self.gui, separators=(",", ":")))
