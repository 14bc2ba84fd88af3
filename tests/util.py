"""Shared helpers of the parity tests: state builders and oracle/CUDA pairing."""
from __future__ import annotations

import numpy as np

import wsb200
from oracle import oracle as O

P = wsb200.params


def stress_state(w: int, h: int, seed: int = 3):
    """A synthetic state that visits every branch of the boundary / advection / lighting passes:
    sea + land terrain from synth.full_state, then urban / industrial / runway / fire / inert
    surface patches, snow, dry desert, smoke plumes, clouds, rain shafts, sub-zero air aloft."""
    g = P.resolve_settings(None)
    g["dayNightCycle"] = False  # sun pinned: zenith angle from guiControls.sunAngle
    g["sunAngle"] = 60.0
    base, water, wall, drops = wsb200.synth.full_state(w, h, seed=seed, g=g, vel_amplitude=0.3)
    rng = np.random.default_rng(seed)
    is_wall = wall[..., 1] == 0
    # repaint the wall type of whole columns (type is vertical in the reference)
    ncol = w
    kinds = rng.choice([1, 1, 1, 2, 3, 4, 5, 6, 0], size=(ncol + 15) // 16)
    col_kind = np.repeat(kinds, 16)[:ncol]
    sea = wall[0, :, 0] == 2
    col_kind = np.where(sea, 2, col_kind)
    col_kind[sea & (rng.random(ncol) < 0.0)] = 2
    wall[..., 0] = col_kind[None, :]
    # fire needs vegetation, desert needs none
    veg = wall[..., 3].astype(np.int32)
    veg = np.where((col_kind == 3)[None, :] & is_wall, 90, veg)
    desert = (rng.random(ncol) < 0.15)[None, :] & is_wall & (col_kind == 1)[None, :]
    veg = np.where(desert, 3, veg)
    wall[..., 3] = veg.astype(np.int8)
    water[..., 2] = np.where(desert, 1.0, water[..., 2])
    # snow on some land
    snowy = (rng.random(ncol) < 0.2)[None, :] & is_wall & (col_kind == 1)[None, :]
    water[..., 3] = np.where(snowy, 12.0, water[..., 3])
    air = ~is_wall
    # clouds / rain / smoke blobs in the air
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(6):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0.2 * h, 0.8 * h), rng.uniform(0.05, 0.15) * h
        blob = np.exp(-(((xx - cx) ** 2 + (yy - cy) ** 2) / (r * r))).astype(np.float32)
        water[..., 1] += np.where(air, 3.0 * blob, 0).astype(np.float32)
        water[..., 0] += np.where(air, 4.0 * blob, 0).astype(np.float32)
        water[..., 2] += np.where(air, 0.5 * blob * (rng.random() < 0.5), 0).astype(np.float32)
    for _ in range(3):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0.05 * h, 0.4 * h), rng.uniform(0.03, 0.08) * h
        blob = np.exp(-(((xx - cx) ** 2 + (yy - cy) ** 2) / (r * r))).astype(np.float32)
        water[..., 3] += np.where(air, 6.0 * blob, 0).astype(np.float32)
    base[..., 3] += np.where(air, rng.normal(0, 0.3, (h, w)), 0).astype(np.float32)
    base[..., 2] += np.where(air, rng.normal(0, 0.01, (h, w)), 0).astype(np.float32)
    # air cells: DISTANCE 127 / VERT 100 / type copied on the first boundary pass
    nd = wsb200.savefile.num_droplets(w, h)
    drops = wsb200.synth.init_rain_drops(nd, seed + 1)
    # a few active droplets of each kind
    k = min(nd, 64)
    drops[:k, 0] = rng.uniform(-1, 1, k)
    drops[:k, 1] = rng.uniform(-0.9, 0.9, k)
    drops[:k, 2] = rng.uniform(0.0, 0.5, k)
    drops[:k, 3] = rng.uniform(0.0, 0.5, k)
    drops[:k, 4] = rng.choice([0.2, 0.6, 1.0], k)
    return g, base, water, wall, drops


def make_oracle(g, base, water, wall, drops, fi=None):
    h, w = base.shape[:2]
    ora = O.OracleSim(w, h, 0 if drops is None else drops.shape[0])
    ora.upload(base, water, wall, drops)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(fi if fi is not None else P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(h, g))
    return ora


def make_cuda(g, base, water, wall, drops, schedule, fi=None, **kw):
    h, w = base.shape[:2]
    sim = wsb200.Simulation(w, h, 0 if drops is None else drops.shape[0], schedule=schedule, gui_controls=g, **kw)
    sim.upload(base, water, wall, drops)
    # same explicit per-frame uniforms as make_oracle (a day/night clock would recompute the angle)
    sim.set_frame_inputs(fi if fi is not None else P.frame_inputs(g))
    return sim


def rel_err(got, want, floor=1e-3):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.max(np.abs(got - want) / (np.abs(want) + floor))) if got.size else 0.0


def ulp_diff(a, b):
    """Largest distance in units-in-the-last-place between two float32 arrays (same-sign values)."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return int(np.max(np.abs(ai - bi))) if ai.size else 0


# ---------------------------------------------------------------------------------------------
# Oracle windows: the oracle on a strip of columns of a large grid (tests at BASELINE sizes)
# ---------------------------------------------------------------------------------------------
STRIP_OWNED = 40        # columns compared per window
STRIP_RADIUS = 8        # columns one iteration can reach (3 boundary chain + 1 + ceil|v| <= 4, cf. strips.py)


def oracle_window(g, fields, W, H, x_first, iters, dry, fi=None):
    """Oracle run on global columns [x_first - margin, x_first + STRIP_OWNED + margin) (periodic);
    returns {name: array[:, STRIP_OWNED, 4]} of the window's owned columns and their global indices."""
    margin = STRIP_RADIUS * iters + 8
    cols = np.arange(x_first - margin, x_first + STRIP_OWNED + margin) % W
    base, water, wall = (f[:, cols] for f in fields)
    ora = O.OracleSim(len(cols), H, 0, global_width=W, x0=x_first - margin)
    ora.upload(base, water, wall, None)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(fi if fi is not None else P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(H, g))
    (ora.step_dry if dry else ora.step)(iters)
    own = slice(margin, margin + STRIP_OWNED)
    out = {"base": ora.field(O.FIELD_BASE, 0)[:, own]}
    if not dry:
        out["water"] = ora.field(O.FIELD_WATER, 1)[:, own]
        out["wall"] = ora.field(O.FIELD_WALL, 0)[:, own]
        out["light"] = ora.light_latest()[:, own]
    ora.close()
    return out, cols[own]


def window_starts(W):
    # the periodic seam (columns W-20 .. W-1, 0 .. 19), the middle (x = W/2 straddles a power of two:
    # fp32 spacing of fragCoord doubles there), the last columns, and an arbitrary tile-straddling one
    return [W - 20, W // 2 - 20, W - STRIP_OWNED, (W // 3) | 1]
