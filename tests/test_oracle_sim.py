"""The oracle against its committed golden vectors, physical invariants (the author's own
conservation diagnostic, app.js:6736-6762), and the x-strip decomposition plan run as a "fake
cluster" of oracle strips with in-memory ghost exchange."""
import hashlib
import os

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from util import make_oracle, stress_state

P = wsb200.params
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _oracle_from_save(sf):
    g = P.resolve_settings(sf.settings_json)
    return make_oracle(g, sf.base, sf.water, sf.wall, sf.droplets), g


def test_oracle_matches_committed_golden(save100):
    gold = np.load(os.path.join(GOLDEN, "oracle_100x100.npz"))
    ora, _ = _oracle_from_save(save100)
    done = 0
    for n in (1, 10, 100):
        ora.step(n - done)
        done = n
        assert np.array_equal(ora.field(O.FIELD_WALL, 0), gold[f"wall_{n}"])
        assert np.array_equal(ora.field(O.FIELD_BASE, 0), gold[f"base_{n}"])
        assert np.array_equal(ora.field(O.FIELD_WATER, 1), gold[f"water_{n}"])
        assert np.array_equal(ora.light_latest(), gold[f"light_{n}"])
        assert np.array_equal(ora.droplets(), gold[f"drops_{n}"])
    ora.step(900)
    h = hashlib.sha256()
    for a in (ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 0), ora.light_latest()):
        h.update(a.tobytes())
    assert np.array_equal(np.frombuffer(h.digest(), np.uint8), gold["sha256_1000"])


def test_invariants_1000_iterations(save100):
    ora, _ = _oracle_from_save(save100)
    w0 = ora.field(O.FIELD_WALL, 0)
    air0 = w0[..., 1] != 0
    water0 = ora.field(O.FIELD_WATER, 0)
    total0 = water0[..., 0][air0].sum(dtype=np.float64)
    ora.step(1000)
    base, water, wall = ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 0)
    assert np.isfinite(base).all() and np.isfinite(water).all() and np.isfinite(ora.light_latest()).all()
    assert (wall[0, :, 1] == 0).all()                     # row 0 stays wall (the lid)
    air = wall[..., 1] != 0
    assert np.array_equal(air, air0)                      # no wall created / destroyed without input
    assert (np.abs(base[..., :2]) < 1.0).all()            # |v| < 1 cell / iteration
    assert (base[..., :2][~air] == 0).all()               # velocities in wall are 0 ... after pressure they stay 0
    assert (water[..., 0][air] >= 0).all() and (water[..., 1][air] >= 0).all()
    assert set(np.unique(water[..., 0][~air])) <= {1001.0, 1002.0}  # wall marker, advectionShader.frag:403-409
    # total water in the air only drifts through the slow sources / sinks (evaporation, globalDrying,
    # precipitation): a few percent over 1000 iterations, not a blow-up
    total1 = water[..., 0][air].sum(dtype=np.float64)
    assert abs(total1 - total0) / total0 < 0.05
    # sunlight has reached the ground after > H iterations
    assert ora.light_latest()[1:, :, 0].max() > 100.0
    # distance fields are the manhattan distance to the nearest wall, saturated at 127
    assert wall[..., 1].min() >= 0 and wall[..., 1].max() <= 127


def test_wall_distance_fields_converge():
    g, base, water, wall, drops = stress_state(96, 64, seed=5)
    ora = make_oracle(g, base, water, wall, None)
    ora.step(130)
    wl = ora.field(O.FIELD_WALL, 0)
    is_wall = wl[..., 1] == 0
    # brute-force manhattan distance with periodic wrap in x and y
    h, w = is_wall.shape
    ys, xs = np.nonzero(is_wall)
    yy, xx = np.mgrid[0:h, 0:w]
    best = np.full((h, w), 10 ** 6)
    for sy, sx in zip(ys, xs):
        dx = np.abs(xx - sx)
        dy = np.abs(yy - sy)
        best = np.minimum(best, np.minimum(dx, w - dx) + np.minimum(dy, h - dy))
    assert np.array_equal(wl[..., 1].astype(int), np.minimum(best, 127))


@pytest.mark.parametrize("n_strips", [2, 4])
def test_fake_cluster_strips_bit_identical(n_strips):
    """N oracle strips with GHOST ghost columns and one exchange per iteration reproduce the
    whole-domain run bit for bit (the decomposition plan libwsb200 uses)."""
    W, H, G = 128, 48, wsb200.strips.GHOST
    g, base, water, wall, _ = stress_state(W, H, seed=11)
    fi = P.frame_inputs(g)
    whole = make_oracle(g, base, water, wall, None, fi)
    parts = []
    for r in range(n_strips):
        cols = wsb200.strips.padded_columns(W, n_strips, r)
        x0, lw = wsb200.strips.strip_bounds(W, n_strips, r)
        o = O.OracleSim(lw + 2 * G, H, 0, global_width=W, x0=x0 - G)
        o.upload(base[:, cols], water[:, cols], wall[:, cols], None)
        o.set_params(P.derive_params(g))
        o.set_frame_inputs(fi)
        o.set_profiles(P.initial_T_profile(H, g))
        parts.append((o, x0, lw))

    fields = [(O.FIELD_BASE, 0), (O.FIELD_BASE, 1), (O.FIELD_WATER, 0), (O.FIELD_WATER, 1), (O.FIELD_WALL, 0),
              (O.FIELD_WALL, 1), (O.FIELD_LIGHT, 0), (O.FIELD_LIGHT, 1)]
    for it in range(12):
        whole.step(1)
        for o, _, _ in parts:
            o.step(1)
        for f, b in fields:
            views = [o.field(f, b, copy=False) for o, _, _ in parts]
            for r, (o, x0, lw) in enumerate(parts):
                left, right = wsb200.strips.neighbours(r, n_strips)
                llw, rlw = parts[left][2], parts[right][2]
                views[r][:, :G] = views[left][:, llw:llw + G]          # left neighbour's rightmost owned
                views[r][:, G + lw:] = views[right][:, G:2 * G]        # right neighbour's leftmost owned
    for f, b in [(O.FIELD_BASE, 0), (O.FIELD_WATER, 1), (O.FIELD_WALL, 0), (O.FIELD_LIGHT, 0), (O.FIELD_LIGHT, 1)]:
        want = whole.field(f, b)
        for o, x0, lw in parts:
            got = o.field(f, b)[:, G:G + lw]
            assert np.array_equal(got, want[:, x0:x0 + lw]), (f, b, x0)
    assert float(np.abs(whole.field(O.FIELD_BASE, 0)[..., :2]).max()) < 1.0


@pytest.mark.parametrize("dry", [False, True])
def test_oracle_window_equals_whole_domain(dry):
    """The windows the full-size GPU tests compare against (util.oracle_window: a strip of columns with a margin of
    STRIP_RADIUS columns per iteration, told its global x offset) reproduce the whole-domain oracle run exactly —
    across the periodic seam too."""
    from util import oracle_window, window_starts

    W, H, iters = 640, 96, 3
    if dry:
        base, water, wall = wsb200.synth.dry_state(W, H, seed=5)
        base[1:, :, 0:2] *= np.float32(20.0)  # |v| up to ~2.5 cells / iteration
        g = P.resolve_settings(None)
    else:
        g, base, water, wall, _ = stress_state(W, H, seed=31)
        g["enablePrecipitation"] = False
    whole = make_oracle(g, base, water, wall, None)
    (whole.step_dry if dry else whole.step)(iters)
    want_all = {"base": whole.field(O.FIELD_BASE, 0), "water": whole.field(O.FIELD_WATER, 1), "wall": whole.field(O.FIELD_WALL, 0),
                "light": whole.light_latest()}
    for x_first in window_starts(W):
        got, cols = oracle_window(g, (base, water, wall), W, H, x_first, iters, dry)
        for name, arr in got.items():
            assert np.array_equal(arr, want_all[name][:, cols], equal_nan=arr.dtype != np.int8), (name, x_first)
