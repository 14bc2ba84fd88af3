"""The fused kernels resolve a semi-Lagrangian back-trace RELATIVE to the own cell when every
velocity component is below 0.9 cells / iteration on a grid of at most 2^19 cells a side
(csrc/wsb_fused_kernels.cuh: near_tap).  This is the arithmetic fact that makes the shortcut
bit-identical to common.glsl:196-199 (st = pos - 0.5; floor(st); fract(st)): in fp32, with
pos = (cell + 0.5) - v [+ 0.05 for the precipitation tap], floor(st) is `cell - 1` exactly when
st < cell and `cell` otherwise."""
import numpy as np

f32 = np.float32


def _check(cells, v, extra):
    cells = cells.astype(np.int64)
    cf = cells.astype(f32)
    frag = cf + f32(0.5)
    pos = frag - v.astype(f32)
    if extra is not None:
        pos = pos + f32(extra)
    st = pos - f32(0.5)
    fl = np.floor(st)
    left = st < cf
    want = np.where(left, cf - f32(1.0), cf)
    assert np.array_equal(fl, want), f"{(fl != want).sum()} of {fl.size} taps leave [cell - 1, cell]"
    # the fraction the kernel computes is the same subtraction
    assert np.array_equal(st - fl, st - want)


def test_near_tap_floor_is_cell_or_cell_minus_one():
    rng = np.random.default_rng(0)
    n = 2_000_000
    for hi in (100, 16384, 1 << 19):
        cells = rng.integers(0, hi, n)
        cells[:8] = [0, 1, hi - 1, hi - 2, hi // 2, 3, 7, hi - 1]
        v = rng.uniform(-0.9, 0.9, n).astype(f32)
        v[:8] = [0.0, -0.0, 0.89999997, -0.89999997, 1e-30, -1e-30, 0.5, -0.5]
        v = np.clip(v, f32(-0.89999997), f32(0.89999997))
        _check(cells, v, None)
        _check(cells, v, 0.05)  # advectionShader.frag:137 — precipitation sampled 0.05 higher


def test_threshold_is_not_slack_beyond_the_grid_cap():
    """At 2^21 cells a side the fp32 spacing is 0.125: the same velocities DO leave the window,
    which is why Geom::nearV switches the shortcut off beyond 2^19."""
    cells = np.full(4096, (1 << 21) + 5, np.int64)
    v = np.linspace(-0.9, 0.9, 4096).astype(f32)
    cf = cells.astype(f32)
    st = ((cf + f32(0.5)) - v + f32(0.05)) - f32(0.5)
    fl = np.floor(st)
    assert ((fl != cf) & (fl != cf - 1)).any()
