"""Seeded synthetic simulation states for the benchmark / parity configurations.

The generators follow the reference's own new-simulation state (shaders/fragment/setupShader.frag
:36-92): air cells start at the initial temperature profile with a dew-point spread of 2 K below
20 % height and 20 K above, wall cells are LAND (soil moisture 25 mm, vegetation, snow above
2000 m) or WATER (25 C), DISTANCE = 255 -> 127 and VERT_DISTANCE = 100 in air so that the boundary
pass rebuilds the distance fields.  Terrain here is a deterministic sum of sines instead of the
shader's sin-hash noise (whose value depends on the GPU's `sin`).
"""
from __future__ import annotations

import numpy as np

from . import params as P
from .savefile import num_droplets


def _max_water(t):
    x = (t / np.float32(250.0)).astype(np.float32)
    x2 = x * x
    x4 = x2 * x2
    x8 = x4 * x4
    return (x8 * x8 * x).astype(np.float32)


def _cols(w, cols):
    """Global column indices to generate: all of them, or the caller's subset (a rank's padded
    strip) — every generator below is a pure function of (global x, y), so strips generated
    separately are identical to slices of the whole grid."""
    return np.arange(w, dtype=np.int64) if cols is None else np.asarray(cols, np.int64)


def _velocity_field(w, h, seed, amplitude, cols=None):
    rng = np.random.default_rng(seed)
    cols = _cols(w, cols)
    xs = ((cols.astype(np.float32) + np.float32(0.5)) / np.float32(w)).astype(np.float32)
    ys = (np.arange(h, dtype=np.float32) + 0.5) / h
    vx = np.zeros((h, len(cols)), np.float32)
    vy = np.zeros((h, len(cols)), np.float32)
    for _ in range(8):  # sum of 8 periodic sines
        kx, ky = int(rng.integers(1, 6)), int(rng.integers(1, 4))
        phx, phy = rng.uniform(0, 2 * np.pi, 2)
        a = np.float32(amplitude / 8.0 * rng.uniform(0.5, 1.5))
        sx = np.sin(2 * np.pi * kx * xs + phx).astype(np.float32)
        sy = np.sin(2 * np.pi * ky * ys + phy).astype(np.float32)
        cx = np.cos(2 * np.pi * kx * xs + phx).astype(np.float32)
        cy = np.cos(2 * np.pi * ky * ys + phy).astype(np.float32)
        vx += a * np.outer(cy, sx)
        vy += a * np.outer(sy, cx)
    return vx, vy


def dry_state(w: int, h: int, seed: int = 1234, g: dict | None = None, cols=None):
    """BASELINE config 2: row 0 LAND wall, rest air; T = initial_T[y] + warm blobs; smooth random
    velocity (amplitude 0.1); P = 0; water = 0.  Returns (base, water, wall); with `cols` only
    those global columns."""
    g = g or P.resolve_settings(None)
    t0 = P.initial_T_profile(h, g)
    rng = np.random.default_rng(seed + 1)
    cols = _cols(w, cols)
    nc = len(cols)
    base = np.zeros((h, nc, 4), np.float32)
    vx, vy = _velocity_field(w, h, seed, 0.1, cols)
    base[..., 0] = vx
    base[..., 1] = vy
    base[..., 3] = t0[:h, None]
    xs = cols.astype(np.float32)[None, :]
    ys = np.arange(h, dtype=np.float32)[:, None]
    for _ in range(16):  # warm / cold blobs, 0.5 K
        cx, cy = rng.uniform(0, w), rng.uniform(0.1 * h, 0.9 * h)
        r = rng.uniform(0.02, 0.08) * h
        dx = np.minimum(np.abs(xs - cx), w - np.abs(xs - cx))
        amp = np.float32(rng.normal(0.0, 0.5))
        base[..., 3] += amp * np.exp(-((dx * dx + (ys - cy) ** 2) / np.float32(r * r))).astype(np.float32)
    water = np.zeros((h, nc, 4), np.float32)
    wall = np.zeros((h, nc, 4), np.int8)
    wall[..., 0] = 1
    wall[..., 1] = np.minimum(np.arange(h), 127)[:, None]
    wall[..., 2] = np.minimum(np.arange(h), 127)[:, None]
    base[0, :, 0:2] = 0.0
    base[0, :, 3] = 1000.0
    water[0, :, 0] = 1001.0
    return base, water, wall


def terrain_height(w: int, h: int, seed: int = 7) -> np.ndarray:
    """Surface row index per column: sea (row 0 only) on ~30 % of the columns, hills up to ~12 % of
    the height elsewhere."""
    rng = np.random.default_rng(seed)
    x = (np.arange(w) + 0.5) / w
    hgt = np.zeros(w)
    for k in (1, 2, 3, 5, 8, 13):
        hgt += np.sin(2 * np.pi * k * x + rng.uniform(0, 2 * np.pi)) / k
    hgt = (hgt - hgt.min()) / (hgt.max() - hgt.min())  # 0..1
    sea = hgt < 0.3
    rows = np.where(sea, 0, 1 + np.floor((hgt - 0.3) / 0.7 * 0.12 * h)).astype(np.int64)
    return rows


def full_state(w: int, h: int, seed: int = 7, g: dict | None = None, with_droplets: bool = True,
               n_droplets: int | None = None, vel_amplitude: float = 0.05, cols=None):
    """BASELINE configs 3-5: terrain + sea, setupShader-style thermodynamic profile, a weak smooth
    wind field so that advection has work to do.  Returns (base, water, wall, droplets); with
    `cols` only those global columns."""
    g = g or P.resolve_settings(None)
    t0 = P.initial_T_profile(h, g)
    lapse = np.float32(P.dry_lapse(g))
    cols = _cols(w, cols)
    nc = len(cols)
    rows = terrain_height(w, h, seed)[cols]
    yy = np.arange(h)[:, None]
    is_wall = yy <= rows[None, :]
    sea = (rows == 0)[None, :]

    base = np.zeros((h, nc, 4), np.float32)
    water = np.zeros((h, nc, 4), np.float32)
    wall = np.zeros((h, nc, 4), np.int8)

    texy = ((np.arange(h, dtype=np.float32) + np.float32(0.5)) * np.float32(1.0 / h))[:, None]
    pot = t0[:h, None].astype(np.float32)
    real = (pot - texy * lapse).astype(np.float32)
    spread = np.where(texy < 0.20, np.float32(2.0), np.float32(20.0)).astype(np.float32)
    total = _max_water(real - spread)
    cloud = np.maximum(total - _max_water(real), np.float32(0.0))
    vx, vy = _velocity_field(w, h, seed + 11, vel_amplitude, cols)

    air = ~is_wall
    base[..., 0] = np.where(air, vx, 0)
    base[..., 1] = np.where(air, vy, 0)
    base[..., 3] = np.where(air, np.broadcast_to(pot, (h, nc)), 0)
    water[..., 0] = np.where(air, np.broadcast_to(total, (h, nc)), 0)
    water[..., 1] = np.where(air, np.broadcast_to(cloud, (h, nc)), 0)

    # wall cells (setupShader.frag:65-78)
    land = is_wall & ~sea
    seaw = is_wall & sea
    base[..., 3] = np.where(seaw, np.float32(P.c_to_k(25.0)), base[..., 3])
    water[..., 2] = np.where(land, np.float32(25.0), water[..., 2])
    xs = cols
    veg = np.clip(60 + 50 * np.sin(2 * np.pi * 3 * (xs + 0.5) / w) - rows * (300.0 / h), 0, 127).astype(np.int8)
    height_m = rows * (g["simHeight"] / h)
    snow = np.clip((height_m - 2000.0) / 3000.0 * 100.0, 0.0, 100.0).astype(np.float32)
    wall[..., 3] = np.where(land, veg[None, :], 0)
    water[..., 3] = np.where(land, snow[None, :], water[..., 3])
    wall[..., 0] = np.where(sea, 2, 1)  # air copies the type below on the first boundary pass anyway
    wall[..., 1] = np.where(is_wall, 0, 127)
    vd = np.clip(yy - rows[None, :], -128, 127)
    wall[..., 2] = np.where(is_wall, vd, 100).astype(np.int8)

    drops = None
    if with_droplets:
        nd = n_droplets if n_droplets is not None else num_droplets(w, h)
        drops = init_rain_drops(nd, seed + 42)
    return base, water, wall, drops


def add_clouds(base, water, wall, n_blobs: int = 64, seed: int = 5, cloud_peak: float = 3.0):
    """Drop Gaussian cloud blobs (cloud water + the same amount of total water) into the air cells
    of a full-size state, between 25 % and 75 % of the height, so that the precipitation pass has
    clouds to spawn droplets in (BASELINE config 4).  In place; each blob only touches its own
    bounding box."""
    h, w = base.shape[:2]
    rng = np.random.default_rng(seed)
    for _ in range(n_blobs):
        cx, cy = rng.uniform(0, w), rng.uniform(0.25 * h, 0.75 * h)
        r = rng.uniform(0.02, 0.06) * h
        x0, x1 = int(max(cx - 3 * r, 0)), int(min(cx + 3 * r, w))
        y0, y1 = int(max(cy - 3 * r, 1)), int(min(cy + 3 * r, h - 2))
        if x1 <= x0 or y1 <= y0:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        blob = (np.float32(cloud_peak) * np.exp(-(((xx - cx) ** 2 + (yy - cy) ** 2) / (r * r)))).astype(np.float32)
        air = wall[y0:y1, x0:x1, 1] != 0
        water[y0:y1, x0:x1, 1] += np.where(air, blob, 0).astype(np.float32)
        water[y0:y1, x0:x1, 0] += np.where(air, blob, 0).astype(np.float32)


# --------------------------------------------------------------------------------------------
# The reference's own new-simulation state: shaders/fragment/setupShader.frag:27-92, which app.js
# renders into both framebuffers when no save file is loaded (app.js:5646-5657, 5729-5742).
# --------------------------------------------------------------------------------------------
def _f32(x):
    return np.asarray(x, np.float32)


def _setup_rand(n):
    """setupShader.frag:27  rand(n) = fract(sin(n) * 43758.5453123).  Canonical form: sin evaluated
    in double and rounded to fp32 (GPU sin of large arguments is implementation-defined), the rest
    in fp32."""
    n = _f32(n)
    v = (np.sin(n.astype(np.float64)).astype(np.float32) * np.float32(43758.5453123)).astype(np.float32)
    return (v - np.floor(v)).astype(np.float32)


def _setup_noise(p):
    """setupShader.frag:29-34"""
    p = _f32(p)
    fl = np.floor(p).astype(np.float32)
    fc = (p - fl).astype(np.float32)
    a, b = _setup_rand(fl), _setup_rand((fl + np.float32(1.0)).astype(np.float32))
    mix = (a * (np.float32(1.0) - fc) + b * fc).astype(np.float32)
    return (mix - np.float32(0.5)).astype(np.float32)


def setup_state(w: int, h: int, seed: float = 0.37, height_mult: float = 0.5, g: dict | None = None,
                with_droplets: bool = True, droplet_seed: int = 42):
    """New-simulation initial state, restating setupShader.frag:36-92 in fp32: terrain from summed
    value noise (all sea below heightMult 0.05, flat land below 0.10), land cells with 25 mm soil
    moisture / vegetation / snow above 2000 m, water cells at 25 C, air at the initial temperature
    profile with a dew point 2 K (lowest 20 %) or 20 K below the temperature.
    `seed` and `height_mult` are the shader uniforms (mouse x / height in the reference).
    Returns (base, water, wall, droplets)."""
    g = g or P.resolve_settings(None)
    f = np.float32
    t0 = P.initial_T_profile(h, g)
    lapse = f(P.dry_lapse(g))
    sim_height = f(g["simHeight"])
    texel_y = f(1.0 / h)
    frag_x = (np.arange(w, dtype=np.float32) + f(0.5))
    frag_y = (np.arange(h, dtype=np.float32) + f(0.5))
    tex_y = (frag_y * texel_y).astype(np.float32)

    height = np.zeros(w, np.float32)
    height_m = np.zeros(w, np.float32)
    hm = f(height_mult)
    if hm < f(0.05):
        pass  # all sea
    elif hm < f(0.10):
        height[:] = f(0.005)  # all land
    else:
        var = (frag_x * f(0.001)).astype(np.float32)
        i = f(2.0)
        while i < f(1000.0):  # :55-57
            off = (_setup_rand((f(seed) + i).astype(np.float32)) * f(10.0)).astype(np.float32)
            height = (height + _setup_noise((var * i + off).astype(np.float32)) * f(0.5) / i).astype(np.float32)
            i = f(i * f(1.5))
        height = (height * hm).astype(np.float32)
        height_m = (height * sim_height).astype(np.float32)

    base = np.zeros((h, w, 4), np.float32)
    water = np.zeros((h, w, 4), np.float32)
    wall32 = np.zeros((h, w, 4), np.int32)
    is_wall = (tex_y[:, None] < texel_y) | (tex_y[:, None] < height[None, :])  # :63
    sea = (height < texel_y)[None, :] & is_wall
    land = is_wall & ~sea
    # wall cells
    wall32[..., 1] = np.where(is_wall, 0, 255)
    wall32[..., 0] = np.where(sea, 2, np.where(land, 1, 0))
    base[..., 3] = np.where(sea, f(25.0) + f(273.15), base[..., 3])
    water[..., 2] = np.where(land, f(25.0), 0)
    veg_noise = _setup_noise((frag_x * f(0.01) + _setup_rand(f(seed)) * f(10.0)).astype(np.float32))
    veg = (f(110.0) - frag_y[:, None] * f(2.0) + veg_noise[None, :] * f(150.0)).astype(np.float32)
    wall32[..., 3] = np.where(land, np.trunc(veg).astype(np.int32), 0)
    mr = (f(0.0) + (height_m - f(2000.0)) * (f(100.0) - f(0.0)) / (f(5000.0) - f(2000.0))).astype(np.float32)  # map_range
    snow = np.maximum(np.minimum(np.maximum(mr, f(0.0)), f(100.0)), f(0.0)).astype(np.float32)
    water[..., 3] = np.where(land, snow[None, :], 0)
    # air cells
    air = ~is_wall
    idx = (tex_y * (f(1.0) / texel_y)).astype(np.int32)
    pot = t0[idx][:, None].astype(np.float32)
    real = (pot - tex_y[:, None] * lapse).astype(np.float32)
    spread = np.where(tex_y[:, None] < f(0.20), f(2.0), f(20.0)).astype(np.float32)
    total = _max_water((real - spread).astype(np.float32))
    cloud = np.maximum(total - _max_water(real), f(0.0)).astype(np.float32)
    base[..., 3] = np.where(air, np.broadcast_to(pot, (h, w)), base[..., 3])
    water[..., 0] = np.where(air, np.broadcast_to(total, (h, w)), 0)
    water[..., 1] = np.where(air, np.broadcast_to(cloud, (h, w)), 0)
    wall32[..., 2] = 100  # :91
    wall = np.clip(wall32, -128, 127).astype(np.int8)  # RGBA8I store saturates
    drops = init_rain_drops(num_droplets(w, h), droplet_seed) if with_droplets else None
    return base, water, wall, drops


def init_rain_drops(n: int, seed: int = 42) -> np.ndarray:
    """initRainDrops (app.js:4901-4913) with a seeded generator instead of Math.random():
    every droplet starts inactive (water mass in [-10,-9)) and the slots hold RNG seeds."""
    rng = np.random.default_rng(seed)
    d = rng.random((n, 5), dtype=np.float32)
    d[:, 2] = np.float32(-10.0) + d[:, 2]
    return d
