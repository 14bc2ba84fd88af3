"""Second, independent restatement of the dry sweep (velocity -> advection of the base field ->
pressure), written in vectorised numpy fp32 directly from the GLSL — velocityShader.frag:32-62,
advectionShader.frag:72-100 + 189-197, common.glsl:194-251 (bilerp / bilerpWall),
pressureShader.frag:16-43 — without looking at the C++ oracle, and the curl / vorticity passes
(curlShader.frag:12-19, vorticityShader.frag:19-38).  The C++ oracle must reproduce it BIT for
bit: every fp32 operation here is a separate numpy ufunc call (one rounding each, no contraction),
which is the oracle's frozen arithmetic.  This pins the oracle against transcription
slips; it cannot pin it against the reference's WebGL output (parity unpinned)."""
import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from util import make_oracle

P = wsb200.params
f32 = np.float32


def _tex(a, dx, dy):
    """texture() with NEAREST + REPEAT in x and y at texel offset (dx, dy)."""
    return np.roll(a, (-dy, -dx), axis=(0, 1))


def _velocity(base, wall, drag, wind):
    b = base.copy()
    p, pxp, pyp = base[..., 2], _tex(base, 1, 0)[..., 2], _tex(base, 0, 1)[..., 2]
    vx = base[..., 0] + (p - pxp)
    vy = base[..., 1] + (p - pyp)
    k = f32(1.0) - f32(drag) * f32(0.0002)
    vx = vx * k
    vy = vy * k
    vx = vx + f32(wind) * f32(0.000001)
    air = wall[..., 1] != 0
    b[..., 0] = np.where(air, vx, f32(0))
    b[..., 1] = np.where(air, vy, f32(0))
    return b


def _mix(a, b, t):
    return a * (f32(1.0) - t) + b * t


def _gather(a, ix, iy):
    h, w = a.shape[:2]
    return a[np.mod(iy, h), np.mod(ix, w)]


def _bilerp(chan, wall, posx, posy, walls):
    stx, sty = posx - f32(0.5), posy - f32(0.5)
    flx, fly = np.floor(stx), np.floor(sty)
    fx, fy = stx - flx, sty - fly
    ix, iy = flx.astype(np.int64), fly.astype(np.int64)
    a, b, c, d = (_gather(chan, ix, iy), _gather(chan, ix + 1, iy), _gather(chan, ix, iy + 1), _gather(chan, ix + 1, iy + 1))
    mab, mcd, mabcd = fx.copy(), fx.copy(), fy.copy()
    if walls:
        dist = wall[..., 1]
        wa, wb, wc, wd = (_gather(dist, ix, iy) == 0, _gather(dist, ix + 1, iy) == 0, _gather(dist, ix, iy + 1) == 0,
                          _gather(dist, ix + 1, iy + 1) == 0)
        mab = np.where(wa, f32(1), np.where(wb, f32(0), mab))
        mcd = np.where(wc, f32(1), np.where(wd, f32(0), mcd))
        mabcd = np.where(wa & wb, f32(1), np.where(wc & wd, f32(0), mabcd))
    return _mix(_mix(a, b, mab), _mix(c, d, mcd), mabcd)


def _advect_base(base, wall):
    h, w = base.shape[:2]
    fx = (np.arange(w, dtype=f32) + f32(0.5))[None, :] * np.ones((h, 1), f32)
    fy = (np.arange(h, dtype=f32) + f32(0.5))[:, None] * np.ones((1, w), f32)
    vx, vy = base[..., 0], base[..., 1]
    vx_xm, vy_ym, vy_xp, vx_yp = _tex(vx, -1, 0), _tex(vy, 0, -1), _tex(vy, 1, 0), _tex(vx, 0, 1)
    vx_xmyp, vy_xpym = _tex(vx, -1, 1), _tex(vy, 1, -1)
    two, four = f32(2), f32(4)
    p_x, p_y = (vx_xm + vx) / two, (vy_ym + vy) / two
    vxx, vxy = vx, (((vy_ym + vy_xp) + vy) + vy_xpym) / four
    vyx, vyy = (((vx_xm + vx_yp) + vx_xmyp) + vx) / four, vy
    out = np.empty_like(base)
    out[..., 0] = _bilerp(vx, wall, fx - vxx, fy - vxy, False)
    out[..., 1] = _bilerp(vy, wall, fx - vyx, fy - vyy, False)
    out[..., 2] = _bilerp(base[..., 2], wall, fx - p_x, fy - p_y, True)
    out[..., 3] = _bilerp(base[..., 3], wall, fx - p_x, fy - p_y, True)
    is_wall = wall[..., 1] == 0
    passthrough = base.copy()
    passthrough[..., 3] = np.where(wall[..., 0] == 1, f32(1000.0), base[..., 3])
    return np.where(is_wall[..., None], passthrough, out)


def _pressure(base, wall):
    b = base.copy()
    wym = _tex(wall, 0, -1)
    land_below = (wym[..., 1] == 0) & (wym[..., 0] == 1)
    b[..., 3] = np.where(land_below, base[..., 3] - (_tex(base, 0, -1)[..., 3] - f32(1000.0)), base[..., 3])
    div = ((_tex(base, -1, 0)[..., 0] - base[..., 0]) + _tex(base, 0, -1)[..., 1]) - base[..., 1]
    b[..., 2] = base[..., 2] + div * f32(0.45)
    return b


def _state(w, h, seed, scale):
    base, water, wall = wsb200.synth.dry_state(w, h, seed=seed)
    base[1:, :, 0:2] *= f32(scale)
    rng = np.random.default_rng(seed)
    for _ in range(10):  # LAND blocks inside the flow (wall-aware interpolation, T = 1000 marker)
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(4, h - 8))
        wall[y0:y0 + 3, x0:x0 + 5, 0] = 1
        wall[y0:y0 + 3, x0:x0 + 5, 1] = 0
        base[y0:y0 + 3, x0:x0 + 5, 0:2] = 0.0
    return base, water, wall


@pytest.mark.parametrize("w,h,scale", [(96, 64, 1.0), (160, 48, 12.0), (77, 50, 30.0)])
def test_dry_sweep_matches_numpy_restatement(w, h, scale):
    base, water, wall = _state(w, h, seed=5, scale=scale)
    g = P.resolve_settings(None)
    g["dragMultiplier"], g["wind"] = 0.003, 2.5
    ora = make_oracle(g, base, water, wall, None)
    p = P.derive_params(g)
    cur = base
    for it in range(6):
        cur = _pressure(_advect_base(_velocity(cur, wall, p.dragMultiplier, p.wind), wall), wall)
        ora.step_dry(1)
        got = ora.field(O.FIELD_BASE, 0)
        assert np.array_equal(got, cur), f"iteration {it + 1}: oracle differs from the numpy restatement in {(got != cur).sum()} values"
    assert np.isfinite(cur).all()


def test_curl_and_vorticity_match_numpy_restatement():
    w, h = 120, 40
    base, water, wall = _state(w, h, seed=9, scale=8.0)
    g = P.resolve_settings(None)
    ora = make_oracle(g, base, water, wall, None)
    ora.run_pass(0)  # velocity -> frameBuff_1
    ora.run_pass(1)  # curl of frameBuff_1
    ora.run_pass(2)  # vorticity force
    p = P.derive_params(g)
    v = _velocity(base, wall, p.dragMultiplier, p.wind)
    vx, vy = v[..., 0], v[..., 1]
    curl = ((_tex(vx, 0, 1) - vx) - _tex(vy, 1, 0)) + vy
    assert np.array_equal(ora.field(O.FIELD_CURL), curl.reshape(h, w, 1)) or np.array_equal(ora.field(O.FIELD_CURL).reshape(h, w), curl)
    ac = np.abs(curl)
    fx = _tex(ac, 0, -1) - _tex(ac, 0, 1)
    fy = _tex(ac, 1, 0) - _tex(ac, -1, 0)
    mag = np.sqrt(fx * fx + fy * fy) + f32(0.0001)
    force = np.stack([(fx / mag) * curl, (fy / mag) * curl], axis=-1)
    assert np.array_equal(ora.field(O.FIELD_VORT).reshape(h, w, 2), force)
