"""Independent numpy fp32 restatement of the FULL advection pass for idle user input
(advectionShader.frag:66-227 + 402-411: air cells — bilerp / bilerpWall gathers of velocity,
pressure, temperature and the four water channels, condensation / evaporation with latent heat,
global drying / heating and sounding forcing, clamp of total water; wall cells — pass-through,
land marker, snow melt and soil evaporation on the surface layer; the wall marker in TOTAL), written
from the GLSL with the frozen closed forms of DESIGN.md 2 (pow(x, 17) as a multiply chain), one
numpy ufunc call per fp32 operation.  The C++ oracle's advection pass must reproduce it bit for bit.
Like test_oracle_numpy_crosscheck.py this pins the oracle against transcription slips, not against
the reference's WebGL output."""
import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from test_oracle_numpy_crosscheck import _bilerp, _tex
from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
LAND, WATER = 1, 2


def _max_water(T):  # common.glsl:175-178, pow(x, 17) = x^16 * x
    x = T / f32(250.0)
    x2 = x * x
    x4 = x2 * x2
    x8 = x4 * x4
    x16 = x8 * x8
    return x16 * x


def _gmax(a, b):  # GLSL max(x, y) = (x < y) ? y : x
    return np.where(a < b, b, a)


def _gmin(a, b):  # GLSL min(x, y) = (y < x) ? y : x
    return np.where(b < a, b, a)


def _clamp(x, lo, hi):
    return _gmin(_gmax(x, lo), hi)


def _map_range(v, min1, max1, min2, max2):
    return min2 + (v - min1) * (max2 - min2) / (max1 - min1)


def _map_range_c(v, min1, max1, min2, max2):
    return _clamp(_map_range(v, min1, max1, min2, max2), _gmin(min2, max2), _gmax(min2, max2))


def _advection(base, water, wall, p, snd_t, snd_w, snd_v, marker=True):
    h, w = base.shape[:2]
    texel_y_uniform = f32(1.0 / h)                     # uniform texelSize (app.js:5436)
    ltexel_y = f32(1.0) / f32(h)                       # :69  texelSize = vec2(1.) / resolution
    fx = (np.arange(w, dtype=f32) + f32(0.5))[None, :] * np.ones((h, 1), f32)
    fy = (np.arange(h, dtype=f32) + f32(0.5))[:, None] * np.ones((1, w), f32)
    tex_y = fy * texel_y_uniform                       # simShader.vert:24
    vx, vy = base[..., 0], base[..., 1]
    two, four = f32(2), f32(4)
    vx_xm, vy_ym, vy_xp, vx_yp = _tex(vx, -1, 0), _tex(vy, 0, -1), _tex(vy, 1, 0), _tex(vx, 0, 1)
    vx_xmyp, vy_xpym = _tex(vx, -1, 1), _tex(vy, 1, -1)
    p_x, p_y = (vx_xm + vx) / two, (vy_ym + vy) / two
    vxy = (((vy_ym + vy_xp) + vy) + vy_xpym) / four
    vyx = (((vx_xm + vx_yp) + vx_xmyp) + vx) / four
    pos_px, pos_py = fx - p_x, fy - p_y

    b = np.empty_like(base)
    wt = np.empty_like(water)
    b[..., 0] = _bilerp(vx, wall, fx - vx, fy - vxy, False)
    b[..., 1] = _bilerp(vy, wall, fx - vyx, fy - vy, False)
    b[..., 2] = _bilerp(base[..., 2], wall, pos_px, pos_py, True)
    b[..., 3] = _bilerp(base[..., 3], wall, pos_px, pos_py, True)
    for ch in (0, 1, 3):
        wt[..., ch] = _bilerp(water[..., ch], wall, pos_px, pos_py, True)
    wt[..., 2] = _bilerp(water[..., 2], wall, pos_px + f32(0.0), pos_py + f32(0.05), True)

    # condensation / evaporation (:111-131)
    real_t = b[..., 3] - tex_y * f32(p.dryLapse)
    over = (wt[..., 0] - _max_water(real_t)) - wt[..., 1]
    cond = np.where(over < 0, over * f32(0.20), over * f32(p.condensationRate))
    cond = _gmax(cond, -wt[..., 1])
    d_t = cond * f32(p.evapHeat) * f32(1.0)
    b[..., 3] = b[..., 3] + d_t
    real_t = real_t + d_t
    wt[..., 1] = wt[..., 1] + cond

    # global effects (:153-176)
    inside = (tex_y > f32(p.globalEffectsStartAlt)) & (tex_y < f32(p.globalEffectsEndAlt))
    lim = _gmax(wt[..., 0] - _max_water(_gmax(real_t - f32(20.0), f32(-80.0) + f32(273.15))), f32(0.0))
    tot = wt[..., 0] - _clamp(np.full_like(lim, f32(p.globalDrying)), f32(0.0), lim)
    temp = b[..., 3] + f32(p.globalHeating)
    idx = (tex_y * (f32(1.0) / ltexel_y)).astype(np.int32)
    im1 = np.maximum(idx - 1, 0)  # index y - 1 clamped at 0

    def snd(a):
        return (a[idx] + a[im1]) / f32(2.0)

    forcing = f32(p.soundingForcing)
    temp = temp - (temp - snd(snd_t)) * f32(0.001) * forcing
    tot = tot - (tot - snd(snd_w)) * f32(0.001) * forcing
    drag = f32(1.0) - _map_range_c(forcing, f32(0.1), f32(1.0), f32(0.0), f32(0.001))
    nvx, nvy = b[..., 0] * drag, b[..., 1] * drag
    nvx = nvx - (nvx - snd(snd_v)) * _map_range_c(forcing, f32(0.9), f32(1.0), f32(0.0), f32(0.001))
    wt[..., 0] = np.where(inside, tot, wt[..., 0])
    b[..., 3] = np.where(inside, temp, b[..., 3])
    b[..., 0] = np.where(inside, nvx, b[..., 0])
    b[..., 1] = np.where(inside, nvy, b[..., 1])
    wt[..., 0] = _gmax(wt[..., 0], f32(0.0))  # :184

    # wall cells (:186-225)
    wl = wall.copy()
    wb, ww = base.copy(), water.copy()
    is_land = wall[..., 0] == LAND
    wb[..., 3] = np.where(is_land, f32(1000.0), wb[..., 3])
    wl[..., 3] = np.maximum(wall[..., 3], 0)
    ww[..., 2] = _gmax(ww[..., 2], f32(0.0))
    above_air = _tex(wall, 0, 1)[..., 1] != 0
    t_above = _tex(base, 0, 1)[..., 3]
    temp_c = (t_above - tex_y * f32(p.dryLapse)) - f32(273.15)
    melt_on = above_air & (ww[..., 3] > 0) & (temp_c > 0)
    melting = _gmin(temp_c * f32(0.000015), ww[..., 3])
    ww[..., 3] = np.where(melt_on, ww[..., 3] - melting, ww[..., 3])
    wb[..., 3] = np.where(melt_on, wb[..., 3] + melting / f32(0.05) * f32(p.meltingHeat), wb[..., 3])
    ww[..., 2] = np.where(melt_on, ww[..., 2] + melting, ww[..., 2])
    evap_on = above_air & (ww[..., 2] > 0) & (temp_c > 0)
    evap = _gmax((_max_water(temp_c + f32(273.15)) - ww[..., 0]) * f32(0.00001), f32(0.0))
    ww[..., 2] = np.where(evap_on, ww[..., 2] - evap, ww[..., 2])
    # :402-411  wall marker in TOTAL (after the user-input block, which may build or remove walls)
    if marker:
        ww[..., 0] = np.where(wall[..., 0] == WATER, f32(1002.0), f32(1001.0))

    is_wall = (wall[..., 1] == 0)[..., None]
    return np.where(is_wall, wb, b), np.where(is_wall, ww, wt), np.where(is_wall, wl, wall)


@pytest.mark.parametrize("forcing", [False, True])
def test_full_advection_pass_matches_numpy_restatement(forcing):
    w, h = 144, 72
    g, base, water, wall, _ = stress_state(w, h, seed=31)
    g["enablePrecipitation"] = False
    if forcing:
        g["globalDrying"], g["globalHeating"], g["soundingForcing"] = 0.00002, 0.0003, 0.95
    p = P.derive_params(g)
    rng = np.random.default_rng(5)
    snd_t = (300.0 + rng.uniform(-5, 5, h + 1)).astype(f32)
    snd_w = rng.uniform(0, 20, h + 1).astype(f32)
    snd_v = rng.uniform(-0.2, 0.2, h + 1).astype(f32)
    ora = make_oracle(g, base, water, wall, None)
    ora.set_profiles(P.initial_T_profile(h, g), snd_t, snd_w, snd_v)
    # the advection pass samples frameBuff_0: upload() fills both copies
    ora.run_pass(4)
    want_b, want_w, want_wl = _advection(base, water, wall, p, snd_t, snd_w, snd_v)
    got_b, got_w, got_wl = ora.field(O.FIELD_BASE, 1), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 1)
    assert np.array_equal(got_wl, want_wl)
    for name, got, want in (("base", got_b, want_b), ("water", got_w, want_w)):
        for ch in range(4):
            bad = got[..., ch] != want[..., ch]
            assert not bad.any(), f"{name}[{ch}]: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got[..., ch][bad][0]!r} vs {want[..., ch][bad][0]!r}"
