"""Independent numpy fp32 restatement of the lighting pass (lightingShader.frag:38-181; the
reflectedLight output is display-only and not part of the simulation state), with the frozen forms
of DESIGN.md 2: the LINEAR fetch of the sun ray is a full-fp32 bilinear in pixel space at
(x + 0.5 + sin a, y + 0.5 + cos a) with wrap S = REPEAT / T = CLAMP_TO_EDGE, sin / cos of the
uniform angle are evaluated in double and rounded, pow(x, .5) = sqrt, pow(x, 4) = (x^2)^2, IR_up is 0
where the shader leaves it uninitialised.  The C++ oracle must reproduce it bit for bit."""
import math

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from test_oracle_numpy_advection import _gmax, _gmin
from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
INERT, LAND, WATER, FIRE, URBAN, RUNWAY, INDUSTRIAL = range(7)
HEAT = f32(0.000002)  # lightHeatingConst


def _mix(a, b, t):
    return a * (f32(1.0) - t) + b * t


def _ir_emitted(T):  # common.glsl:256-259
    t = T * f32(0.01)
    t2 = t * t
    return (t2 * t2) * f32(5.670374419)


def _lighting(base, water, wall, light, p, fi):
    h, w = base.shape[:2]
    hf = f32(h)
    fx = (np.arange(w, dtype=f32) + f32(0.5))[None, :] * np.ones((h, 1), f32)
    fy = (np.arange(h, dtype=f32) + f32(0.5))[:, None] * np.ones((1, w), f32)
    tex_y = fy * f32(1.0 / h)
    sin_s, cos_s = f32(math.sin(float(fi.sunAngle))), f32(math.cos(float(fi.sunAngle)))
    comp = f32(300.0) / hf

    # sun ray: bilinear fetch of SUNLIGHT
    stx, sty = (fx + sin_s) - f32(0.5), (fy + cos_s) - f32(0.5)
    flx, fly = np.floor(stx), np.floor(sty)
    tx, ty = stx - flx, sty - fly
    ix, iy = flx.astype(np.int64), fly.astype(np.int64)
    y0, y1 = np.clip(iy, 0, h - 1), np.clip(iy + 1, 0, h - 1)
    x0, x1 = np.mod(ix, w), np.mod(ix + 1, w)
    sun = light[..., 0]
    sunlight = _mix(_mix(sun[y0, x0], sun[y0, x1], tx), _mix(sun[y1, x0], sun[y1, x1], tx), ty)

    real_t = base[..., 3] - tex_y * f32(p.dryLapse)
    tot, cloud, precip, smoke = (water[..., k] for k in range(4))
    wtype, wdist, wvert = wall[..., 0], wall[..., 1], wall[..., 2]

    # air cells
    shaded = fy < hf - f32(2.0)
    refl = _gmin(np.sqrt(cloud * f32(0.0010) + precip * f32(0.00020)) * comp, f32(1.0)) + f32(0.0002)
    absb = _gmin(smoke * f32(0.020) * comp, f32(1.0))
    l_refl, l_abs = sunlight * refl, sunlight * absb
    sun_air = np.where(shaded, _gmax(f32(0.0), sunlight - l_refl - l_abs), sunlight)
    net = np.where(shaded, f32(0.0) + l_abs * HEAT, f32(0.0))

    rows = np.arange(h)
    ir_down_in = light[np.minimum(rows + 1, h - 1), :, 2]   # CLAMP_TO_EDGE in T
    ir_up_below = light[np.maximum(rows - 1, 0), :, 3]

    # one above the surface (:90-116)
    t_below = np.roll(base[..., 3], 1, axis=0)               # base texture: REPEAT
    solid = (wtype == RUNWAY) | (wtype == URBAN) | (wtype == INDUSTRIAL) | (wtype == LAND)
    ir_up_surf = np.where(solid, _ir_emitted(real_t), np.where(wtype == WATER, _ir_emitted(t_below),
                          np.where(wtype == FIRE, _ir_emitted(real_t + f32(100.0)), f32(0.0))))
    net_surf = np.where(solid | (wtype == WATER), net + (ir_down_in - ir_up_surf) * HEAT, np.where(wtype == FIRE, f32(0.0), net))

    # in the air (:117-144)
    emis = f32(p.greenhouseGases) + tot * f32(p.waterGreenHouseEffect)
    emis = emis + cloud * f32(5.0)
    emis = _gmin(emis * comp, f32(1.0))
    a_down, a_up = ir_down_in * emis, ir_up_below * emis
    emitted = _ir_emitted(real_t) * emis
    net_airc = net + ((a_down + a_up) - emitted * f32(2.0)) * HEAT
    ir_down_air = (ir_down_in - a_down) + emitted
    ir_up_air = (ir_up_below - a_up) + emitted

    surf = wvert == 1
    net_out = np.where(surf, net_surf, net_airc) * f32(p.IR_rate)
    air = np.stack([sun_air, net_out, np.where(surf, ir_down_in, ir_down_air), np.where(surf, ir_up_surf, ir_up_air)], axis=-1)

    # wall cells (:164-178)
    zero = np.zeros_like(sunlight)
    wl = np.stack([np.where(wtype == WATER, sunlight * f32(0.90), f32(0.0)), zero, zero, zero], axis=-1)
    out = np.where((wdist != 0)[..., None], air, wl)
    top = fy >= hf - f32(1.0)
    out[top] = np.array([fi.sunIntensity, 0, 0, 0], f32)
    return out


@pytest.mark.parametrize("angle_deg", [9.9, 60.0, 90.0, 133.0])  # GUI degrees: 90 = sun overhead
def test_lighting_pass_matches_numpy_restatement(angle_deg):
    w, h = 144, 72
    g, base, water, wall, _ = stress_state(w, h, seed=41)
    g["enablePrecipitation"] = False
    g["dayNightCycle"] = False
    g["sunAngle"] = angle_deg
    p = P.derive_params(g)
    fi = P.frame_inputs(g)
    ora = make_oracle(g, base, water, wall, None, fi=fi)
    # a few whole iterations give a non-trivial light field in both ping-pong copies
    ora.step(3)
    even = ora.even
    src = 0 if even else 1
    light_src = ora.field(O.FIELD_LIGHT, src)
    b1, w1, wl1 = ora.field(O.FIELD_BASE, 1), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 1)
    want = _lighting(b1, w1, wl1, light_src, p, fi)
    ora.run_pass(6)
    got = ora.field(O.FIELD_LIGHT, 1 - src)
    for ch, name in enumerate(("sunlight", "net heating", "IR down", "IR up")):
        bad = got[..., ch] != want[..., ch]
        assert not bad.any(), f"{name}: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got[..., ch][bad][0]!r} vs {want[..., ch][bad][0]!r}"
    assert np.isfinite(got).all() and got[..., 0].max() > 0
