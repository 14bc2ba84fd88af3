"""New-simulation state generator (setupShader.frag:27-92 restated in synth.setup_state) and the
oracle running from it."""
import numpy as np

import wsb200
from oracle import oracle as O

from util import make_oracle

P = wsb200.params


def test_all_sea_and_all_land():
    g = P.resolve_settings(None)
    base, water, wall, drops = wsb200.synth.setup_state(128, 100, seed=0.2, height_mult=0.01, g=g)
    assert (wall[0, :, 1] == 0).all() and (wall[0, :, 0] == 2).all()          # one row of sea
    assert (wall[1:, :, 1] == 127).all()                                        # 255 saturates in RGBA8I
    assert np.allclose(base[0, :, 3], 298.15)
    assert (wall[..., 2] == 100).all()
    assert drops.shape == (128 * 100 // 25, 5) and (drops[:, 2] < 0).all()
    # "all land" is height 0.005: land only where a cell is thinner than that (H > 200), like the shader
    base, water, wall, _ = wsb200.synth.setup_state(128, 300, seed=0.2, height_mult=0.07, g=g)
    assert (wall[0, :, 0] == 1).all() and (water[0, :, 2] == 25.0).all()       # flat land, 25 mm soil moisture
    assert (wall[2, :, 1] != 0).all()                                           # air above texCoord.y = 0.005


def test_hills_profile_and_thermodynamics():
    g = P.resolve_settings(None)
    w, h = 400, 150
    base, water, wall, _ = wsb200.synth.setup_state(w, h, seed=0.37, height_mult=0.5, g=g)
    is_wall = wall[..., 1] == 0
    assert is_wall[0].all()
    # terrain is a height field: walls are contiguous from the bottom
    assert (np.diff(is_wall.astype(int), axis=0) <= 0).all()
    heights = is_wall.sum(0)
    assert heights.max() > 2 and heights.min() >= 1 and len(set(heights)) > 3
    air = ~is_wall
    t0 = P.initial_T_profile(h, g)
    ys = np.nonzero(air[:, 0])[0]
    assert np.array_equal(base[ys, 0, 3], t0[ys])
    # dew point spread: relative humidity is higher in the lowest 20 % than above
    tex_y = (np.arange(h) + 0.5) / h
    real = base[..., 3] - tex_y[:, None] * 120.0
    sat = wsb200.synth._max_water(real.astype(np.float32))
    with np.errstate(invalid="ignore", divide="ignore"):
        rh = np.where(air, water[..., 0] / sat, np.nan)
    assert np.nanmean(rh[tex_y < 0.2]) > 0.8 > np.nanmean(rh[tex_y > 0.25])
    assert (water[..., 1][air] == 0).all()  # sub-saturated everywhere: no cloud water yet
    # deterministic
    again = wsb200.synth.setup_state(w, h, seed=0.37, height_mult=0.5, g=g)
    assert np.array_equal(again[0], base) and np.array_equal(again[2], wall)
    other = wsb200.synth.setup_state(w, h, seed=0.38, height_mult=0.5, g=g)
    assert not np.array_equal(other[2], wall)


def test_oracle_runs_from_setup_state():
    g = P.resolve_settings(None)
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    base, water, wall, drops = wsb200.synth.setup_state(160, 100, seed=0.61, height_mult=0.8, g=g)
    ora = make_oracle(g, base, water, wall, drops)
    ora.step(150)
    b, wl = ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WALL, 0)
    assert np.isfinite(b).all()
    # the boundary pass has rebuilt the distance fields from the 127 / 100 presets
    air = wl[..., 1] != 0
    assert wl[..., 1][air].min() == 1 and wl[..., 2][air].min() == 1
    assert (np.abs(b[..., :2]) < 1.0).all()
