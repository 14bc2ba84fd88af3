"""Independent restatement of the user-input block and the airplane block of the advection pass
(advectionShader.frag:229-457: brush tools 1-4, wall / surface tools 10-22 with positive and
negative intensity, whole-width and circular brushes with and without horizontal wrap, the wall
marker in TOTAL, water dump and crash of the airplane) as a scalar Python transliteration applied
on top of the numpy restatement of the pass's first half (test_oracle_numpy_advection.py).  The
C++ oracle's advection pass must reproduce it bit for bit, including the RGBA8I saturation of the
wall texel."""
import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from test_oracle_numpy_advection import _advection
from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
Z, ONE = f32(0.0), f32(1.0)
INERT, LAND, WATER, FIRE, URBAN, RUNWAY, INDUSTRIAL = range(7)


def gmax(a, b):
    return b if a < b else a


def gmin(a, b):
    return b if b < a else a


def clamp(x, lo, hi):
    return gmin(gmax(x, lo), hi)


def abs_horizontal_dist(a, b):
    return gmin(gmin(abs(a - b), abs(ONE + a - b)), ONE - a + b)


def length(x, y):
    return np.sqrt(x * x + y * y)


def user_input_and_airplane(base, water, wall, wall_in, fi, p, initial_t):
    """base / water / wall: the pass's results before the user-input block (modified in place);
    wall_in: the wall texture the pass samples."""
    h, w = base.shape[:2]
    texel_ux, texel_uy = f32(1.0 / w), f32(1.0 / h)      # uniform texelSize -> texCoord
    lt_x, lt_y = ONE / f32(w), ONE / f32(h)              # :69 texelSize = vec2(1.) / resolution
    ui = [f32(v) for v in fi.userInputValues]
    move = [f32(v) for v in fi.userInputMove]
    plane = [f32(v) for v in fi.airplaneValues]
    kind = int(fi.userInputType)
    wrap = bool(fi.wrapHorizontally)
    out_wall = np.empty_like(wall)
    for y in range(h):
        ty = (f32(y) + f32(0.5)) * texel_uy
        for x in range(w):
            tx = (f32(x) + f32(0.5)) * texel_ux
            b = [f32(v) for v in base[y, x]]
            wt = [f32(v) for v in water[y, x]]
            wl = [int(v) for v in wall[y, x]]
            above_dist = int(wall_in[(y + 1) % h, x, 1])
            in_brush, weight = False, ONE
            if ui[0] < f32(-0.5):
                if abs(ui[1] - ty) < ui[3] * lt_y:
                    in_brush = True
            else:
                vx = abs_horizontal_dist(ui[0], tx) if wrap else abs(ui[0] - tx)
                vy = ui[1] - ty
                vx = vx * (lt_y / lt_x)
                dist = length(vx, vy)
                e0 = ui[3] * lt_y
                with np.errstate(divide="ignore", invalid="ignore"):  # brush size 0 (idle): weight is never used
                    t = clamp((dist - e0) / (Z - e0), Z, ONE)
                weight = t * t * (f32(3.0) - f32(2.0) * t)
                if dist < ui[3] * lt_y:
                    in_brush = True
            if in_brush:
                if kind == 1:
                    b[3] = b[3] + ui[2]
                    if wl[0] == 2 and wl[1] == 0:
                        b[3] = clamp(b[3], Z + f32(273.15), f32(40.0) + f32(273.15))
                elif kind == 2:
                    if wt[1] > Z:
                        wt[1] = wt[1] + ui[2]
                        wt[1] = gmax(wt[1], Z)
                    wt[0] = wt[0] + ui[2]
                    wt[0] = gmax(wt[0], Z)
                elif kind == 3 and wl[1] != 0:
                    wt[3] = wt[3] + ui[2]
                    wt[3] = gmin(gmax(wt[3], Z), f32(2.0))
                elif kind == 4:
                    if ui[0] < f32(-0.5):
                        b[0] = b[0] + move[0] * f32(5.0) * weight * ui[2]
                    else:
                        b[0] = b[0] + move[0] * f32(5.0) * weight * ui[2]
                        b[1] = b[1] + move[1] * f32(5.0) * weight * ui[2]
                elif kind >= 10:
                    surface = wl[1] == 0 and above_dist != 0
                    if ui[2] > Z:
                        set_wall = False
                        if kind == 10:
                            wl[0], set_wall = INERT, True
                        elif kind == 11:
                            wl[0], set_wall = LAND, True
                        elif kind == 12:
                            wl[0], set_wall = WATER, True
                        elif kind == 13:
                            if surface and wl[0] == LAND:
                                wl[0], set_wall = FIRE, True
                        elif kind == 14:
                            if surface and wl[0] in (LAND, RUNWAY, INDUSTRIAL):
                                wl[0] = URBAN
                        elif kind == 15:
                            if surface and wl[0] in (LAND, URBAN, INDUSTRIAL):
                                wl[0] = RUNWAY
                        elif kind == 16:
                            if surface and wl[0] in (LAND, URBAN, RUNWAY):
                                wl[0] = INDUSTRIAL
                        elif kind == 20:
                            if surface and wl[0] != WATER:
                                wt[2] = wt[2] + ui[2] * f32(10.0)
                        elif kind == 21:
                            if surface and wl[0] in (LAND, URBAN, INDUSTRIAL):
                                wt[3] = wt[3] + ui[2] * f32(0.5)
                        elif kind == 22:
                            if surface and wl[0] in (LAND, FIRE, URBAN, INDUSTRIAL):
                                wl[3] += 1
                        if set_wall:
                            wl[1] = 0
                            b[3] = f32(1000.0)
                            if wl[0] == LAND:
                                wt[2] = f32(25.0)
                            elif wl[0] == WATER:
                                b[3] = f32(p.waterTemperature)
                    else:
                        if wl[1] == 0:
                            if kind == 13:
                                if wl[0] == FIRE:
                                    wl[0] = LAND
                            elif kind == 14:
                                if wl[0] == URBAN:
                                    wl[0] = LAND
                            elif kind == 15:
                                if wl[0] == RUNWAY:
                                    wl[0] = LAND
                            elif kind == 16:
                                if wl[0] == INDUSTRIAL:
                                    wl[0] = LAND
                            elif kind == 20:
                                wt[2] = wt[2] + ui[2] * f32(10.0)
                            elif kind == 21:
                                wt[3] = wt[3] + ui[2] * f32(0.5)
                            elif kind == 22:
                                wl[3] = max(wl[3] - 1, 0)
                            elif ty > lt_y:
                                wl[1] = 255
                                b[0], b[1], b[2] = Z, Z, Z
                                b[3] = f32(initial_t[int(ty * (ONE / lt_y))])
                                wt[0], wt[1], wt[2], wt[3] = Z, Z, Z, Z
            if wl[1] == 0:
                wt[0] = f32(1002.0) if wl[0] == WATER else f32(1001.0)
            # airplane (:413-457)
            px = abs_horizontal_dist(plane[0], tx) if wrap else abs(plane[0] - tx)
            py = plane[1] - ty
            px = px * (lt_y / lt_x)
            px, py = px * f32(h), py * f32(h)
            if plane[3] < Z:
                px, py = px + Z, py + f32(-1.0)
            dist = length(px, py)
            influence = gmax(ONE - dist, Z) * f32(0.03)
            if plane[3] < Z:
                wt[2] = wt[2] + influence * f32(100.0)
            if plane[3] > f32(0.9):
                if dist < f32(1.5):
                    if wl[1] == 0:
                        if wl[0] == LAND and wl[2] == 0:
                            wl[0] = FIRE
                    else:
                        b[2] = b[2] + f32(0.05)
                        b[3] = f32(50.0) + f32(273.15)
                        wt[0] = wt[0] + ONE
                        wt[3] = wt[3] + f32(10.0)
            base[y, x] = b
            water[y, x] = wt
            out_wall[y, x] = [max(-128, min(127, v)) for v in wl]
    return base, water, out_wall


CASES = [  # type, intensity, (x, y, size), move, wrap, airplane
    (1, 0.7, (0.31, 0.12, 9.0), (0, 0), 0, (0, 0, 0, 0)),
    (1, -0.4, (-1.0, 0.03, 3.0), (0, 0), 0, (0, 0, 0, 0)),       # whole-width brush over the sea surface
    (2, 0.25, (0.62, 0.55, 12.0), (0, 0), 1, (0, 0, 0, 0)),
    (2, -3.0, (0.98, 0.5, 10.0), (0, 0), 1, (0, 0, 0, 0)),       # wraps around the edge
    (3, 1.5, (0.4, 0.3, 8.0), (0, 0), 0, (0.4, 0.32, 0.5, -1.0)),  # smoke + airplane dumping water
    (4, 0.8, (0.5, 0.4, 14.0), (0.02, -0.01), 0, (0, 0, 0, 0)),
    (4, 0.8, (-1.0, 0.4, 5.0), (0.02, -0.01), 0, (0, 0, 0, 0)),
    (10, 1.0, (0.2, 0.5, 6.0), (0, 0), 0, (0, 0, 0, 0)),
    (11, 1.0, (0.45, 0.35, 7.0), (0, 0), 0, (0, 0, 0, 0)),
    (12, 1.0, (0.7, 0.2, 7.0), (0, 0), 0, (0, 0, 0, 0)),
    (11, -1.0, (0.3, 0.06, 9.0), (0, 0), 0, (0, 0, 0, 0)),       # remove walls
    (13, 1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (13, -1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (14, 1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (14, -1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (15, 1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (15, -1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (16, 1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (16, -1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (20, 0.6, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (20, -0.6, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (21, 0.6, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (21, -0.6, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (22, 1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (22, -1.0, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)),
    (0, 0.0, (0.5, 0.5, 5.0), (0, 0), 1, (0.52, 0.3, 1.0, 1.0)),   # crash in the air
    (0, 0.0, (0.5, 0.5, 5.0), (0, 0), 0, (0.30, 0.10, 1.0, 1.0)),  # crash near / into the ground
]


@pytest.mark.parametrize("case", CASES, ids=[f"type{c[0]}{'+' if c[1] >= 0 else '-'}{i}" for i, c in enumerate(CASES)])
def test_user_input_and_airplane_match_python_transliteration(case):
    kind, intensity, (bx, by, size), move, wrap, plane = case
    w, h = 80, 48
    g, base, water, wall, _ = stress_state(w, h, seed=29)
    g["enablePrecipitation"] = False
    # every surface type the tools can convert or revert
    sea = wall[0, :, 0] == WATER
    for x0, t in ((0, RUNWAY), (12, INDUSTRIAL), (24, URBAN), (36, FIRE), (48, LAND)):
        cols = np.arange(x0, x0 + 12)
        cols = cols[~sea[cols]]
        wall[:, cols, 0] = t
    p = P.derive_params(g)
    fi = P.frame_inputs(g)
    fi.userInputType = kind
    for k, v in enumerate((bx, by, intensity, size)):
        fi.userInputValues[k] = v
    fi.userInputMove[0], fi.userInputMove[1] = move
    fi.wrapHorizontally = wrap
    for k, v in enumerate(plane):
        fi.airplaneValues[k] = v
    initial_t = P.initial_T_profile(h, g)
    zeros = np.zeros(h + 1, f32)
    ora = make_oracle(g, base, water, wall, None, fi=fi)
    # vertical distances / vegetation settle for a few iterations with idle input first
    idle = P.frame_inputs(g)
    ora.set_frame_inputs(idle)
    ora.step(3)
    ora.set_frame_inputs(fi)
    for k in range(4):  # velocity, curl, vorticity, boundary -> frameBuff_0, which advection samples
        ora.run_pass(k)
    b0, w0, wl0 = ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 0), ora.field(O.FIELD_WALL, 0)
    if plane[3] > 0.9 and plane[1] < 0.2:  # "into the ground": aim at a LAND surface cell (VERT_DISTANCE 0)
        ys, xs = np.nonzero((wl0[..., 1] == 0) & (wl0[..., 0] == LAND) & (wl0[..., 2] == 0))
        assert len(ys) > 0
        fi.airplaneValues[0], fi.airplaneValues[1] = (xs[0] + 0.5) / w, (ys[0] + 1.1) / h
        ora.set_frame_inputs(fi)
    pre = _advection(b0, w0, wl0, p, zeros, zeros, zeros, marker=False)
    want_b, want_w, want_wl = user_input_and_airplane(pre[0].copy(), pre[1].copy(), pre[2].copy(), wl0, fi, p, initial_t)
    ora.run_pass(4)
    got_b, got_w, got_wl = ora.field(O.FIELD_BASE, 1), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 1)
    bad = (got_wl != want_wl).any(axis=-1)
    assert not bad.any(), f"wall: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got_wl[bad][0]} vs {want_wl[bad][0]} (in {wl0[bad][0]})"
    for name, got, want in (("base", got_b, want_b), ("water", got_w, want_w)):
        for ch in range(4):
            bad = got[..., ch] != want[..., ch]
            assert not bad.any(), (f"{name}[{ch}]: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got[..., ch][bad][0]!r} vs "
                                   f"{want[..., ch][bad][0]!r} (wall in {wl0[bad][0]})")
    # the tool must have done something
    idle_b, idle_w, idle_wl = user_input_and_airplane(pre[0].copy(), pre[1].copy(), pre[2].copy(), wl0, idle, p, initial_t)
    assert not (np.array_equal(idle_b, want_b) and np.array_equal(idle_w, want_w) and np.array_equal(idle_wl, want_wl)), "case had no effect"
