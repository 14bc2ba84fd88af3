"""The product's per-cell bodies (csrc/wsb_cells.cuh, wsb_math.cuh, GlobalCtx / GlobalAt of
wsb_ref_kernels.cuh) compiled for the HOST (tests/host_cells/, g++ -ffp-contract=off) and driven
through the reference schedule must reproduce the oracle bit for bit — on the CPU, without a GPU.
What is left for the GPU parity tests to establish is then only the tile plumbing of the kernels
(staging, halos, indices), not the cell arithmetic.  Test infrastructure: nothing in the product
loads tests/host_cells/libhostcells.so."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from util import make_oracle, stress_state

P = wsb200.params
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_cells")


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostCells:
    def __init__(self, L, g, base, water, wall, fi=None, profiles=None):
        self.L, (self.h_, self.w_) = L, base.shape[:2]
        self.h = L.hc_create(self.w_, self.h_)
        b, w, wl = (np.ascontiguousarray(a) for a in (base, water, wall))
        L.hc_upload(self.h, _ptr(b), _ptr(w), _ptr(wl))
        p = P.derive_params(g)
        L.hc_set_params(self.h, ctypes.byref(p))
        fi = fi if fi is not None else P.frame_inputs(g)
        L.hc_set_frame_inputs(self.h, ctypes.byref(fi))
        prof = [np.ascontiguousarray(a, np.float32) for a in (profiles or (P.initial_T_profile(self.h_, g),))]
        prof += [np.zeros(self.h_ + 1, np.float32)] * (4 - len(prof))
        L.hc_set_profiles(self.h, *(_ptr(a) for a in prof))

    def run(self, p):
        self.L.hc_run_pass(self.h, p)

    def read(self, field, buf):
        if field == 2:
            out = np.empty((self.h_, self.w_, 4), np.int8)
        else:
            out = np.empty((self.h_, self.w_, 4), np.float32)
        self.L.hc_read(self.h, field, buf, _ptr(out))
        return out

    def close(self):
        self.L.hc_destroy(self.h)


FIELDS = (("base", 0, O.FIELD_BASE), ("water", 1, O.FIELD_WATER), ("wall", 2, O.FIELD_WALL), ("light", 3, O.FIELD_LIGHT))


def _compare(hc, ora, what):
    for name, f, of in FIELDS:
        for buf in (0, 1):
            got, want = hc.read(f, buf), ora.field(of, buf)
            same = (got == want) | ((got != got) & (want != want)) if got.dtype != np.int8 else got == want
            assert same.all(), f"{what}: {name}_{buf} differs in {(~same).sum()} values, first at {np.argwhere(~same)[0]}"


@pytest.mark.parametrize("shape,seed", [((96, 48), 3), ((133, 61), 11)])
def test_cell_headers_on_the_host_reproduce_the_oracle(lib, shape, seed):
    w, h = shape
    g, base, water, wall, _ = stress_state(w, h, seed=seed)
    g["enablePrecipitation"] = False
    g["globalDrying"], g["globalHeating"], g["soundingForcing"] = 0.00002, 0.0003, 0.95
    rng = np.random.default_rng(seed)
    prof = (P.initial_T_profile(h, g), (300 + rng.uniform(-5, 5, h + 1)).astype(np.float32), rng.uniform(0, 20, h + 1).astype(np.float32),
            rng.uniform(-0.2, 0.2, h + 1).astype(np.float32))
    ora = make_oracle(g, base, water, wall, None)
    ora.set_profiles(*prof)
    hc = HostCells(lib, g, base, water, wall, profiles=prof)
    for it in range(25):
        for p in (0, 1, 2, 3, 4, 5, 6, 8):
            ora.run_pass(p)
            hc.run(p)
            if it in (0, 24):
                _compare(hc, ora, f"iteration {it}, pass {p}")
    _compare(hc, ora, "after 25 iterations")
    assert np.isfinite(hc.read(0, 0)).all()
    hc.close()


def test_cell_headers_user_input_and_slow_processes(lib):
    """Brush + airplane inputs and the iteration-multiple processes of the boundary pass."""
    w, h = 80, 40
    g, base, water, wall, _ = stress_state(w, h, seed=5)
    g["enablePrecipitation"] = False
    for kind, intensity, it0 in ((1, 0.7, 98), (2, 0.3, 199), (4, 0.8, 19), (11, 1.0, 0), (12, -1.0, 599), (22, 1.0, 39), (0, 0.0, 100)):
        fi = P.frame_inputs(g)
        fi.userInputType = kind
        for k, v in enumerate((0.4, 0.2, intensity, 9.0)):
            fi.userInputValues[k] = v
        fi.userInputMove[0], fi.userInputMove[1] = 0.02, -0.01
        for k, v in enumerate((0.6, 0.3, 1.0, -1.0 if kind != 0 else 1.0)):
            fi.airplaneValues[k] = v
        ora = make_oracle(g, base, water, wall, None, fi=fi)
        hc = HostCells(lib, g, base, water, wall, fi=fi)
        ora.iter = it0
        lib.hc_set_iter(hc.h, it0)
        for it in range(3):
            for p in (0, 1, 2, 3, 4, 5, 6, 8):
                ora.run_pass(p)
                hc.run(p)
        _compare(hc, ora, f"tool {kind} from iteration {it0}")
        hc.close()


# ---------------------------------------------------------------------------------------------
# the fused kernels themselves, on a host emulation of the CUDA execution model
# ---------------------------------------------------------------------------------------------
class EmuFused:
    """The FUSED schedule on the emulator, with the read-back views of wsb_read_rect."""

    def __init__(self, L, g, base, water, wall, fi=None, profiles=None, strip=None):
        """strip = (global width, first owned column, owned columns, ghost columns): base / water / wall are then
        the strip's LOCAL arrays, ghost columns included (wsb_upload_local)."""
        self.L, (self.h_, self.w_) = L, base.shape[:2]
        self.h = L.ef_create(self.w_, self.h_) if strip is None else L.ef_create_strip(strip[0], self.h_, strip[1], strip[2], strip[3])
        b, w, wl = (np.ascontiguousarray(a) for a in (base, water, wall))
        L.ef_upload(self.h, _ptr(b), _ptr(w), _ptr(wl))
        p = P.derive_params(g)
        L.ef_set_params(self.h, ctypes.byref(p))
        fi = fi if fi is not None else P.frame_inputs(g)
        L.ef_set_frame_inputs(self.h, ctypes.byref(fi))
        prof = [np.ascontiguousarray(a, np.float32) for a in (profiles or (P.initial_T_profile(self.h_, g),))]
        prof += [np.zeros(self.h_ + 1, np.float32)] * (4 - len(prof))
        L.ef_set_profiles(self.h, *(_ptr(a) for a in prof))

    def read(self, field, view):
        out = np.empty((self.h_, self.w_, 4), np.int8 if field == 2 else np.float32)
        self.L.ef_read(self.h, field, view, _ptr(out))
        return out

    def close(self):
        self.L.ef_destroy(self.h)


def _compare_fused(em, ora, what):
    """Canonical state after whole iterations (tests/test_gpu_parity.py: _assert_fields_equal)."""
    pairs = (("base", em.read(0, 0), ora.field(O.FIELD_BASE, 0)), ("water1", em.read(1, 1), ora.field(O.FIELD_WATER, 1)),
             ("water0", em.read(1, 0), ora.field(O.FIELD_WATER, 0)), ("wall", em.read(2, 0), ora.field(O.FIELD_WALL, 0)),
             ("light", em.read(3, 2), ora.light_latest()), ("light0", em.read(3, 0), ora.field(O.FIELD_LIGHT, 0)))
    for name, got, want in pairs:
        same = got == want
        assert same.all(), f"{what}: {name} differs in {(~same).sum()} values, first at {np.argwhere(~same)[0]}: {got[~same][0]!r} vs {want[~same][0]!r}"


@pytest.mark.parametrize("shape,seed", [((192, 96), 7), ((100, 100), 3), ((333, 77), 5), ((256, 40), 9)])
def test_fused_kernels_on_the_emulator_reproduce_the_oracle(emu, shape, seed):
    """Full physics, grids with and without TMA-capable tiles, ragged edges, walls in the flow."""
    w, h = shape
    g, base, water, wall, _ = stress_state(w, h, seed=seed)
    g["enablePrecipitation"] = False
    ora = make_oracle(g, base, water, wall, None)
    em = EmuFused(emu, g, base, water, wall)
    assert bool(emu.ef_uses_tma(em.h)) == (w % 4 == 0 and w >= 72 and h >= 32)
    done = 0
    for n in (1, 2, 6):
        ora.step(n - done)
        emu.ef_step(em.h, n - done)
        done = n
        _compare_fused(em, ora, f"{shape} after {n} iterations")
    assert 0 < emu.ef_max_velocity(em.h) < 1.0
    em.close()


@pytest.mark.parametrize("kind,intensity", [(1, 0.7), (2, -0.5), (3, 1.5), (4, 0.8), (10, 1.0), (11, 1.0), (12, 1.0), (11, -1.0), (13, 1.0), (14, 1.0),
                                            (16, -1.0), (20, 0.6), (21, 0.6), (22, 1.0), (22, -1.0), (0, 0.0)])
def test_fused_kernels_on_the_emulator_with_tools_and_airplane(emu, kind, intensity):
    """The user-input and airplane blocks inside k_fused_adv (wall texels that change, saturate or pass
    through; the water-surface look-ahead of the lighting half) against the oracle."""
    w, h = 192, 64
    g, base, water, wall, _ = stress_state(w, h, seed=29)
    g["enablePrecipitation"] = False
    sea = wall[0, :, 0] == 2
    for x0, t in ((0, 5), (24, 6), (48, 4), (72, 3), (96, 1)):  # every surface type the tools convert or revert
        cols = np.arange(x0, x0 + 24)
        wall[:, cols[~sea[cols]], 0] = t
    fi = P.frame_inputs(g)
    fi.userInputType = kind
    for k, v in enumerate((-1.0 if kind >= 13 else 0.4, 0.12 if kind >= 10 else 0.3, intensity, 10.0)):
        fi.userInputValues[k] = v
    fi.userInputMove[0], fi.userInputMove[1] = 0.02, -0.01
    for k, v in enumerate((0.6, 0.2, 1.0, 1.0 if kind == 0 else (-1.0 if kind == 3 else 0.0))):
        fi.airplaneValues[k] = v
    ora = make_oracle(g, base, water, wall, None, fi=fi)
    em = EmuFused(emu, g, base, water, wall, fi=fi)
    for n in (1, 3):
        ora.step(n - (0 if n == 1 else 1))
        emu.ef_step(em.h, n - (0 if n == 1 else 1))
        if not np.isfinite(ora.field(O.FIELD_BASE, 0)).all():
            # a tool that blows the flow up (the reference's "NaN bug"): parity is only defined for finite states —
            # the kernels skip updates whose uniform rate is exactly 0, and inf * 0 is not inf - 0
            assert not np.isfinite(em.read(0, 0)).all()
            break
        _compare_fused(em, ora, f"tool {kind} ({intensity:+}) after {n} iterations")
    em.close()


@pytest.mark.parametrize("shape,scale", [((512, 128), 1.0), ((384, 96), 15.0), ((256, 64), 25.0), ((150, 61), 8.0), ((203, 77), 10.0)])
def test_fused_dry_sweep_on_the_emulator_reproduces_the_oracle(emu, shape, scale):
    """Slow, moderate (near back-trace in every direction) and fast flow (hand-over to the exact
    path), wall blocks in the flow, TMA-staged and register-staged tiles, 64 x 28 tiles on ragged
    grids."""
    w, h = shape
    base, water, wall = wsb200.synth.dry_state(w, h, seed=11)
    base[1:, :, 0:2] *= np.float32(scale)
    rng = np.random.default_rng(3)
    for _ in range(12):
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(4, h - 8))
        wall[y0:y0 + 3, x0:x0 + 5, 0] = 1
        wall[y0:y0 + 3, x0:x0 + 5, 1] = 0
        base[y0:y0 + 3, x0:x0 + 5, 0:2] = 0.0
    g = P.resolve_settings(None)
    ora = make_oracle(g, base, water, wall, None)
    em = EmuFused(emu, g, base, water, wall)
    done = 0
    for n in (1, 4):
        ora.step_dry(n - done)
        emu.ef_step_dry(em.h, n - done)
        done = n
        got, want = em.read(0, 0), ora.field(O.FIELD_BASE, 0)
        assert np.array_equal(got, want), f"{shape} x{scale} after {n}: {(got != want).sum()} values differ, first at {np.argwhere(got != want)[0]}"
    vmax = emu.ef_max_velocity(em.h)
    assert (vmax > 1.0) == (scale >= 20.0)
    em.close()


@pytest.mark.parametrize("ranks,width", [(2, 256), (3, 336), (4, 512)])
def test_strip_partition_on_the_emulator_is_bit_identical(emu, ranks, width):
    """The multi-GPU plan of csrc/wsb200.cu (x-strips with 8 ghost columns per side, clamped instead of periodic
    columns, one exchange of the 13 planes per iteration) executed with one emulated simulation per rank: every owned
    cell must equal the oracle's single-domain run, for strip widths that are and are not multiples of the tile."""
    h, ghost, iters = 64, 8, 6
    g, base, water, wall, _ = stress_state(width, h, seed=23)
    g["enablePrecipitation"] = False
    ora = make_oracle(g, base, water, wall, None)
    ora.step(iters)
    bounds = [(r * width) // ranks for r in range(ranks + 1)]
    sims = []
    for r in range(ranks):
        x0, lw = bounds[r], bounds[r + 1] - bounds[r]
        cols = np.arange(x0 - ghost, x0 + lw + ghost) % width
        sims.append(EmuFused(emu, g, base[:, cols], water[:, cols], wall[:, cols], strip=(width, x0, lw, ghost)))

    def planes(sim):
        ptrs = (ctypes.c_void_p * 13)()
        emu.ef_exchange_planes(sim.h, ptrs)
        return [np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(h, sim.w_)) for p in ptrs]

    for _ in range(iters):
        for sim in sims:
            emu.ef_step(sim.h, 1)
        pl = [planes(sim) for sim in sims]
        for r in range(ranks):
            left, right = pl[(r - 1) % ranks], pl[(r + 1) % ranks]
            lw_left = sims[(r - 1) % ranks].w_ - 2 * ghost
            lw = sims[r].w_ - 2 * ghost
            for k in range(13):
                pl[r][k][:, :ghost] = left[k][:, lw_left:lw_left + ghost]    # the left neighbour's rightmost owned columns
                pl[r][k][:, ghost + lw:] = right[k][:, ghost:2 * ghost]       # the right neighbour's leftmost owned columns
    views = (("base", 0, 0, ora.field(O.FIELD_BASE, 0)), ("water1", 1, 1, ora.field(O.FIELD_WATER, 1)), ("wall", 2, 0, ora.field(O.FIELD_WALL, 0)),
             ("light", 3, 2, ora.light_latest()))
    for name, f, v, want in views:
        got = np.concatenate([sim.read(f, v)[:, ghost:sim.w_ - ghost] for sim in sims], axis=1)
        same = got == want
        assert same.all(), f"{ranks} strips: {name} differs in {(~same).sum()} values, first at {np.argwhere(~same)[0]}"
    for sim in sims:
        sim.close()


def test_particle_pass_on_the_emulator(emu):
    """k_precipitation (one origin add per droplet, inactive count per CTA), k_boxsum / k_clear_origins (sprites as a
    12 x 12 box filter over the dirty tiles) and k_latch on the emulator: droplet records bit-exact after the first pass, the additive feedback / deposition
    textures equal up to fp32 summation order, the fields within the north-star tolerance after several iterations
    (the same bars as tests/test_gpu_parity.py)."""
    w, h = 128, 64
    g, base, water, wall, drops = stress_state(w, h, seed=3)
    drops = np.ascontiguousarray(drops[:700])
    g["enablePrecipitation"] = True
    air = wall[..., 1] != 0
    water[h // 2:, :, 1] += np.where(air[h // 2:], np.float32(20.0), np.float32(0.0))  # spawning becomes likely
    water[h // 2:, :, 0] += np.where(air[h // 2:], np.float32(20.0), np.float32(0.0))
    ora = make_oracle(g, base, water, wall, drops)
    em = EmuFused(emu, g, base, water, wall)
    emu.ef_upload_drops(em.h, _ptr(drops), drops.shape[0])
    ora.inactive_droplets = 5.0
    emu.ef_set_inactive(em.h, 5.0)

    def state():
        d = np.empty_like(drops)
        emu.ef_read_drops(em.h, _ptr(d))
        fb, dep = np.empty((h, w, 4), np.float32), np.empty((h, w, 2), np.float32)
        emu.ef_read_feedback(em.h, _ptr(fb), _ptr(dep))
        return d, fb, dep

    ora.step(1)
    emu.ef_step(em.h, 1)
    d, fb, dep = state()
    assert np.array_equal(d, ora.droplets()), "droplet records after the first pass"
    assert np.allclose(fb, ora.field(O.FIELD_FEEDBACK), rtol=1e-5, atol=1e-6)
    assert np.allclose(dep, ora.field(O.FIELD_DEPOSITION), rtol=1e-5, atol=1e-6)
    assert (d[:, 2] >= 0).sum() > 50 and fb[0, 0, 0] > 0  # active droplets and the inactive count
    ora.step(5)
    emu.ef_step(em.h, 5)
    d, fb, dep = state()
    want = ora.droplets()
    assert np.array_equal(d[:, 2] < 0, want[:, 2] < 0)  # the same droplets are active
    assert np.allclose(d, want, rtol=1e-4, atol=1e-6)
    for name, f, v, of, ob in (("base", 0, 0, O.FIELD_BASE, 0), ("water1", 1, 1, O.FIELD_WATER, 1)):
        got, ref = em.read(f, v).astype(np.float64), ora.field(of, ob).astype(np.float64)
        err = np.max(np.abs(got - ref) / (np.abs(ref) + 1e-3))
        assert err < 1e-5, f"{name}: relative error {err:.3g} after 6 iterations with particles"
    assert np.array_equal(em.read(2, 0), ora.field(O.FIELD_WALL, 0))
    lightning, inactive = np.zeros(4, np.float32), ctypes.c_float()
    emu.ef_get_latches(em.h, _ptr(lightning), ctypes.byref(inactive))
    assert np.array_equal(lightning, ora.lightning) and inactive.value == ora.inactive_droplets
    em.close()


@pytest.mark.parametrize("shape", [(128, 64), (150, 61), (64, 48)])
def test_sprite_box_filter_borders_and_overlaps(emu, shape):
    """Sprites as a box filter of their origins against the oracle's pixel-by-pixel rasterisation: droplets on and
    next to every border of the viewport (clipped, never wrapped; a centre outside the clip volume is discarded),
    on tile seams, hundreds overlapping in one spot, droplets in the ground (deposition) — and the origin grid /
    dirty map must be clean again afterwards (a second pass with all droplets gone leaves only the inactive count)."""
    w, h = shape
    g, base, water, wall, _ = stress_state(w, h, seed=9)
    g["enablePrecipitation"] = True
    rng = np.random.default_rng(5)
    xs = np.concatenate([[-1.0, -0.9999, -1.0 + 11.9 / w, 1.0, 0.9999, 1.0 - 11.9 / w, 0.0, 2.0 * 64 / w - 1.0, 2.0 * 63.5 / w - 1.0],
                         rng.uniform(-1, 1, 300), np.full(200, 0.31), rng.uniform(-1, -0.9, 40), rng.uniform(0.9, 1, 40)])
    ys = np.concatenate([[0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.999, 2.0 * 16 / h - 1.0, 2.0 * 15.5 / h - 1.0],
                         rng.uniform(-0.99, 0.999, 300), np.full(200, 0.4) + rng.uniform(-0.05, 0.05, 200), rng.uniform(0.9, 0.999, 40),
                         rng.uniform(-0.999, -0.9, 40)])
    n = len(xs)
    drops = np.zeros((n, 5), np.float32)
    drops[:, 0], drops[:, 1] = xs, ys
    drops[:, 2] = rng.uniform(0.05, 0.5, n)
    drops[:, 3] = rng.uniform(0.0, 0.5, n)
    drops[:, 4] = rng.choice([0.2, 0.6, 1.0], n)
    ora = make_oracle(g, base, water, wall, drops)
    em = EmuFused(emu, g, base, water, wall)
    emu.ef_upload_drops(em.h, _ptr(drops), n)
    ora.step(1)
    emu.ef_step(em.h, 1)
    d = np.empty_like(drops)
    emu.ef_read_drops(em.h, _ptr(d))
    fb, dep = np.empty((h, w, 4), np.float32), np.empty((h, w, 2), np.float32)
    emu.ef_read_feedback(em.h, _ptr(fb), _ptr(dep))
    assert np.array_equal(d, ora.droplets())
    want_fb, want_dep = ora.field(O.FIELD_FEEDBACK), ora.field(O.FIELD_DEPOSITION)
    assert np.array_equal(fb != 0, want_fb != 0) and np.array_equal(dep != 0, want_dep != 0)   # the same texels are covered
    assert np.allclose(fb, want_fb, rtol=1e-5, atol=1e-7) and np.allclose(dep, want_dep, rtol=1e-5, atol=1e-7)
    assert (want_dep != 0).any() and (want_fb[:, 0] != 0).any() and (want_fb[:, -1] != 0).any() and (want_fb[-1] != 0).any()
    assert np.abs(want_fb[..., 0]).max() > 50 * np.abs(want_fb[..., 0][want_fb[..., 0] != 0]).min()   # a pile of overlapping sprites
    dirty = ctypes.c_int()
    assert emu.ef_origin_residue(em.h, ctypes.byref(dirty)) == 0 and dirty.value > 0   # origins cleared; tiles await the boundary kernel
    emu.ef_step(em.h, 1)                                                                 # ... which consumes and clears them, before the next pass flags its own
    # second pass: deactivate every droplet; feedback must then hold the inactive count and nothing else
    ora.step(1)
    gone = ora.droplets().copy()
    gone[:, 2] = -5.0
    ora.droplets(copy=False)[...] = gone
    emu.ef_upload_drops(em.h, _ptr(gone), n)   # also clears feedback / deposition like the reference's gl.clear
    p = P.derive_params(g)
    p.spawnChanceMult = 0.0
    ora.set_params(p)
    emu.ef_set_params(em.h, ctypes.byref(p))
    ora.field(O.FIELD_FEEDBACK, copy=False)[...] = 0
    ora.field(O.FIELD_DEPOSITION, copy=False)[...] = 0
    ora.run_pass(7)
    emu.ef_step(em.h, 1)
    emu.ef_read_feedback(em.h, _ptr(fb), _ptr(dep))
    assert fb[0, 0, 0] == n and np.count_nonzero(fb) == 1 and not dep.any()


def test_fused_kernels_on_the_emulator_forcing_and_slow_processes(emu):
    """Global drying / heating / sounding forcing (the uniform-rate blocks the kernels skip when a rate is exactly 0)
    and the processes that run at multiples of 20, 100 and 600 iterations, on the fused kernels."""
    w, h = 192, 64
    g, base, water, wall, _ = stress_state(w, h, seed=13)
    g["enablePrecipitation"] = False
    rng = np.random.default_rng(13)
    prof = (P.initial_T_profile(h, g), (300 + rng.uniform(-5, 5, h + 1)).astype(np.float32), rng.uniform(0, 20, h + 1).astype(np.float32),
            rng.uniform(-0.2, 0.2, h + 1).astype(np.float32))
    for drying, heating, forcing, it0 in ((0.00002, 0.0003, 0.95, 97), (0.0, 0.0003, 0.0, 598), (0.00002, 0.0, 0.5, 18)):
        g["globalDrying"], g["globalHeating"], g["soundingForcing"] = drying, heating, forcing
        ora = make_oracle(g, base, water, wall, None)
        ora.set_profiles(*prof)
        em = EmuFused(emu, g, base, water, wall, profiles=prof)
        ora.iter = it0
        emu.ef_set_iter(em.h, it0)
        ora.step(5)
        emu.ef_step(em.h, 5)
        _compare_fused(em, ora, f"forcing ({drying}, {heating}, {forcing}) from iteration {it0}")
        em.close()
