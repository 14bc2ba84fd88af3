"""New-simulation state generator (setupShader.frag:27-92 restated in synth.setup_state) and the
oracle running from it."""
import numpy as np
import pytest

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


# ---------------------------------------------------------------------------------------------
# A second, independent restatement of setupShader.frag:27-92: ONE FRAGMENT AT A TIME, np.float32
# scalars, control flow as in the shader (the product's synth.setup_state is vectorised over columns /
# rows and shares nothing with it but the canonical forms of DESIGN 2: sin of the hash noise evaluated
# in double and rounded to fp32, pow(x, 17) as the multiply chain x2, x4, x8, x16 * x, mix = a(1-t)+bt,
# RGBA8I stores saturate).
# ---------------------------------------------------------------------------------------------
import math  # noqa: E402

f32 = np.float32


def _rand(n):                       # :27  fract(sin(n) * 43758.5453123)
    v = f32(f32(math.sin(float(f32(n)))) * f32(43758.5453123))
    return f32(v - f32(math.floor(float(v))))


def _noise(p):                      # :29-34
    p = f32(p)
    fl = f32(math.floor(float(p)))
    fc = f32(p - fl)
    a, b = _rand(fl), _rand(f32(fl + f32(1.0)))
    return f32(f32(f32(a * f32(f32(1.0) - fc)) + f32(b * fc)) - f32(0.5))


def _max_water(t):                  # common.glsl:177  pow(T / 250, 17)
    x = f32(f32(t) / f32(250.0))
    x2 = f32(x * x)
    x4 = f32(x2 * x2)
    x8 = f32(x4 * x4)
    return f32(f32(x8 * x8) * x)


def _setup_fragment(ix, iy, w, h, seed, height_mult, sim_height, dry_lapse, t0):
    """main() of setupShader.frag for the fragment at cell (ix, iy): returns (base[4], water[4], wall[4])."""
    frag_x, frag_y = f32(f32(ix) + f32(0.5)), f32(f32(iy) + f32(0.5))   # simShader.vert:21-35 (exact x + 0.5)
    texel_y = f32(1.0 / h)
    tex_y = f32(frag_y * texel_y)
    base, water, wall = [f32(0)] * 4, [f32(0)] * 4, [0, 0, 0, 0]
    height, height_m = f32(0.0), f32(0.0)
    hm = f32(height_mult)
    if hm < f32(0.05):
        height = f32(0.0)
    elif hm < f32(0.10):
        height = f32(0.005)
    else:
        var = f32(frag_x * f32(0.001))
        i = f32(2.0)
        while i < f32(1000.0):
            n = _noise(f32(f32(var * i) + f32(_rand(f32(f32(seed) + i)) * f32(10.0))))
            height = f32(height + f32(f32(n * f32(0.5)) / i))
            i = f32(i * f32(1.5))
        height = f32(height * hm)
        height_m = f32(height * f32(sim_height))
    if tex_y < texel_y or tex_y < height:
        wall[1] = 0
        if height < texel_y:
            wall[0] = 2
            base[3] = f32(f32(25.0) + f32(273.15))
        else:
            wall[0] = 1
            water[2] = f32(25.0)
            veg = f32(f32(f32(110.0) - f32(frag_y * f32(2.0))) + f32(_noise(f32(f32(frag_x * f32(0.01)) + f32(_rand(f32(seed)) * f32(10.0)))) * f32(150.0)))
            wall[3] = int(veg)      # int(): truncation towards zero
            # map_rangeC(height_m, 2000, 5000, 0, 100) = clamp(map_range(...), 0, 100)   common.glsl:209-217
            mr = f32(f32(0.0) + f32(f32(f32(height_m - f32(2000.0)) * f32(f32(100.0) - f32(0.0))) / f32(f32(5000.0) - f32(2000.0))))
            water[3] = max(min(max(mr, f32(0.0)), f32(100.0)), f32(0.0))
    else:
        wall[1] = 255
        base[3] = f32(t0[int(f32(tex_y * f32(f32(1.0) / texel_y)))])
        real = f32(base[3] - f32(tex_y * f32(dry_lapse)))
        water[0] = _max_water(f32(real - f32(2.0))) if tex_y < f32(0.20) else _max_water(f32(real - f32(20.0)))
        water[1] = max(f32(water[0] - _max_water(real)), f32(0.0))
    wall[2] = 100
    return base, water, [max(-128, min(127, v)) for v in wall]


@pytest.mark.parametrize("w,h,seed,height_mult", [(96, 64, 0.37, 0.5), (80, 120, 0.61, 0.8), (64, 48, 0.9, 1.0), (50, 40, 0.2, 0.01), (50, 240, 0.2, 0.07)])
def test_setup_state_matches_fragment_by_fragment_restatement(w, h, seed, height_mult):
    g = P.resolve_settings(None)
    base, water, wall, _ = wsb200.synth.setup_state(w, h, seed=seed, height_mult=height_mult, g=g, with_droplets=False)
    t0 = P.initial_T_profile(h, g)
    lapse = f32(P.dry_lapse(g))
    n_wall = 0
    for iy in range(h):
        for ix in range(w):
            b, wt, wl = _setup_fragment(ix, iy, w, h, seed, height_mult, g["simHeight"], lapse, t0)
            assert [int(v) for v in wall[iy, ix]] == wl, (ix, iy, wall[iy, ix], wl)
            assert all(f32(p) == f32(q) for p, q in zip(base[iy, ix], b)), (ix, iy, base[iy, ix], b)
            assert all(f32(p) == f32(q) for p, q in zip(water[iy, ix], wt)), (ix, iy, water[iy, ix], wt)
            n_wall += wl[1] == 0
    assert n_wall >= w  # at least the bottom row
