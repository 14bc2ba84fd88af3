"""Golden vectors produced by the REFERENCE'S OWN SHADERS (tests/golden/ref_shader_golden.npz, generated in a container
that has the reference checkout by tests/golden/make_ref_shader_golden.py through oracle/_ref — the reference's GLSL
compiled for the host): the oracle (CPU tests) and the CUDA path (`-m gpu`) must reproduce them.

Bars.  base / water / wall / droplet records: BIT-EXACT wherever the summation order is fixed — always for the
oracle, and for the CUDA path whenever no particle sprite has fed back into the fluid (all `*_dry` / `stress` /
`hotlake` cases, every iteration; with particles the sprite sums are fp32 atomics in arbitrary order, compared at the
north star's 1e-5).  light: bit-exact on the power-of-two grids (`stress*`, `hotlake`); on the 100 x 100 grid the
SUNLIGHT channel travels through the LINEAR fetch of lightingShader.frag:48-49, whose sample position the shader
forms in normalised coordinates and the frozen semantics in pixel space — identical when 1 / size is a
power of two, an ulp of the row number apart otherwise (real hardware quantises it to 8 bits): compared at 1e-4 of the
solar constant; the three IR / heating channels are bit-exact there too."""
import os
import sys

import numpy as np
import pytest

import wsb200
from oracle import oracle as O
from oracle import ref_shaders as R
from util import make_cuda, make_oracle, rel_err, ulp_diff

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_ref_shader_golden as G  # noqa: E402

P = wsb200.params
SIM = wsb200.sim
REL = 1e-5  # north_star tolerance on fp32 fields


@pytest.fixture(scope="module")
def gold():
    return np.load(G.OUT)


def check_light(got, want, pow2, what):
    if pow2:
        assert np.array_equal(got, want), f"{what}: light not bit-exact, max ulp {ulp_diff(got, want)}"
        return
    assert np.array_equal(got[..., 1:], want[..., 1:]), f"{what}: NET_HEATING / IR_DOWN / IR_UP not bit-exact"
    sun = max(float(np.abs(want[..., 0]).max()), 1.0)
    assert np.abs(got[..., 0] - want[..., 0]).max() <= 1e-4 * sun, f"{what}: SUNLIGHT differs by {np.abs(got[..., 0] - want[..., 0]).max():.3g} of {sun:.4g}"


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_the_reference_shader_vectors(name, gold):
    g, base, water, wall, drops = G.case_inputs(name)
    rain = drops is not None and g["enablePrecipitation"]
    ora = make_oracle(g, base, water, wall, drops if rain else None)
    h, w = base.shape[:2]
    pow2 = (w & (w - 1)) == 0 and (h & (h - 1)) == 0
    done = 0
    for n in G.CASES[name]:
        ora.step(n - done)
        done = n
        what = f"{name} after {n} iterations"
        assert np.array_equal(ora.field(O.FIELD_WALL, 0), gold[f"{name}/wall/{n}"]), f"{what}: wall"
        for f, b, key in ((O.FIELD_BASE, 0, "base"), (O.FIELD_WATER, 1, "water")):
            got, want = ora.field(f, b), gold[f"{name}/{key}/{n}"]
            assert np.array_equal(got, want), f"{what}: {key} not bit-exact, max ulp {ulp_diff(got, want)}"
        if n < 1000:
            check_light(ora.light_latest(), gold[f"{name}/light/{n}"], pow2, what)
        if rain:
            assert np.array_equal(ora.droplets(), gold[f"{name}/drops/{n}"]), f"{what}: droplet records"
            assert np.array_equal(ora.lightning, gold[f"{name}/lightning/{n}"]), f"{what}: lightning record"


@pytest.mark.skipif(not R.available(), reason="needs the reference checkout")
def test_committed_vectors_are_what_the_reference_shaders_produce_today(gold):
    """Regenerate from the reference checkout and compare: the committed file is not stale."""
    fresh = G.run()
    assert sorted(fresh) == sorted(gold.files)
    for k, v in fresh.items():
        assert np.array_equal(v, gold[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [SIM.SCHEDULE_REFERENCE, SIM.SCHEDULE_FUSED])
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_reproduces_the_reference_shader_vectors(name, schedule, gold):
    g, base, water, wall, drops = G.case_inputs(name)
    rain = drops is not None and g["enablePrecipitation"]
    sim = make_cuda(g, base, water, wall, drops if rain else None, schedule)
    h, w = base.shape[:2]
    pow2 = (w & (w - 1)) == 0 and (h & (h - 1)) == 0
    done = 0
    for n in G.CASES[name]:
        sim.step(n - done)
        done = n
        what = f"{name} after {n} iterations"
        wall_got = sim.read_pixels(SIM.FIELD_WALL)
        pairs = (("base", sim.read_pixels(SIM.FIELD_BASE, view=SIM.VIEW_FRAMEBUFF_0)), ("water", sim.read_pixels(SIM.FIELD_WATER, view=SIM.VIEW_FRAMEBUFF_1)))
        light = sim.read_pixels(SIM.FIELD_LIGHT, view=SIM.VIEW_LATEST)
        if not rain:  # no sprite sums: the arithmetic order is fixed, so is every bit
            assert np.array_equal(wall_got, gold[f"{name}/wall/{n}"]), f"{what}: wall differs in {(wall_got != gold[f'{name}/wall/{n}']).sum()} bytes"
            for key, got in pairs:
                want = gold[f"{name}/{key}/{n}"]
                assert np.array_equal(got, want), f"{what}: {key} not bit-exact, max ulp {ulp_diff(got, want)}, rel {rel_err(got, want):.3g}"
            if n < 1000:
                check_light(light, gold[f"{name}/light/{n}"], pow2, what)
            continue
        # with particles: sprite sums in atomic order; the weather amplifies the last bit slowly
        if n <= 100:
            assert np.array_equal(wall_got, gold[f"{name}/wall/{n}"]), f"{what}: wall"
        for key, got in pairs:
            want = gold[f"{name}/{key}/{n}"]
            if n == 1:
                assert np.array_equal(got, want), f"{what}: {key} not bit-exact (no sprite has fed back yet)"
            elif n <= 100:
                assert rel_err(got, want) < REL, f"{what}: {key} rel err {rel_err(got, want):.3g}"
        d_got, d_want = sim.read_droplets(), gold[f"{name}/drops/{n}"]
        if n == 1:
            assert np.array_equal(d_got, d_want), f"{what}: droplet records"
        elif n <= 100:
            assert np.array_equal(d_got[:, 2] < 0, d_want[:, 2] < 0), f"{what}: different droplets active"
            assert np.allclose(d_got, d_want, rtol=1e-4, atol=1e-6), f"{what}: droplet records"
    sim.close()


@pytest.mark.parametrize("name,snaps", [("stress", (1, 10)), ("save100_dry", (1, 10)), ("hotlake", (100,))])
def test_fused_kernels_on_the_emulator_reproduce_the_reference_shader_vectors(name, snaps, gold, emu):
    """csrc/wsb_fused_kernels.cuh compiled unchanged for the host emulation of the CUDA execution model
    (tests/host_cells/): the product's kernels against the reference's shaders without a GPU, bit for bit."""
    from test_host_cells import EmuFused

    g, base, water, wall, _ = G.case_inputs(name)
    em = EmuFused(emu, g, base, water, wall)
    h, w = base.shape[:2]
    pow2 = (w & (w - 1)) == 0 and (h & (h - 1)) == 0
    done = 0
    for n in snaps:
        emu.ef_step(em.h, n - done)
        done = n
        what = f"{name} after {n} iterations (emulator)"
        assert np.array_equal(em.read(2, 0), gold[f"{name}/wall/{n}"]), f"{what}: wall"
        for key, f, v in (("base", 0, 0), ("water", 1, 1)):
            got, want = em.read(f, v), gold[f"{name}/{key}/{n}"]
            assert np.array_equal(got, want), f"{what}: {key} not bit-exact, max ulp {ulp_diff(got, want)}"
        check_light(em.read(3, 2), gold[f"{name}/light/{n}"], pow2, what)
    em.close()
