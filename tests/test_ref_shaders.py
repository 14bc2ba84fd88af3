"""The oracle against the REFERENCE'S OWN SHADER CODE, executed.

oracle/_ref/libref_shaders.so (oracle/ref_shim/) is the reference's GLSL — common.glsl, simShader.vert, the nine
simulation fragment shaders, precipitationShader.vert, setupShader.frag — translated mechanically from the files
where they lie in the reference checkout and compiled for the host; only the JavaScript around them (texture
objects, bindings and the draw loop of app.js:5830-6005) and the fixed-function parts of GL are restated.  These
tests hold oracle/wsb_oracle.cpp (and synth.setup_state) to it BIT FOR BIT:

  * every pass of the loop, on identical inputs, on each of the 14 shipped saves (all their wall types, fires,
    snow, droplets), for two iterations;
  * whole runs of up to 300 iterations on power-of-two crops of shipped saves and on synthetic stress states —
    particles, spawning, lightning, every brush / wall tool, the airplane, sounding forcing, the slow processes at
    iteration multiples, the `% 0` growth rates;
  * the setup / terrain generator.

One quantity is compared with a tolerance on grids whose 1 / size is not a power of two: SUNLIGHT, which travels
through the hardware LINEAR fetch of lightingShader.frag:48-49.  The shader forms `texCoord + sunRay` in normalised
coordinates, the oracle works in pixel space; the two round the sample position differently (by an ulp
of the row / column number, ~3e-5 of a texel) unless the texel size is exact.  Real hardware quantises that weight
to 8 bits (4e-3 of a texel), so neither form is "the" WebGL result; on power-of-two grids they are identical — at
every sun angle, tested — and the runs below are bit-exact in every field including the light.

Mutation-checked: six seeded one-token changes of the oracle — a coefficient in the boundary, advection and lighting
passes and in the precipitation feedback, the wrong neighbour in the pressure pass, an off-by-one in the wall-distance
propagation — each fail the first save they touch.

Needs the reference checkout (this container); skipped on the GPU box, where tests/golden/ref_shader_*.npz —
generated from this library by tests/golden/make_ref_shader_golden.py — stand in (tests/test_ref_shader_golden.py)."""
import glob
import os

import numpy as np
import pytest

import wsb200
from oracle import oracle as O
from oracle import ref_shaders as R
from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
REFERENCE_SAVES = os.environ.get("WSB_REFERENCE_SAVES", "/root/reference/saves")
pytestmark = pytest.mark.skipif(not R.available(), reason="needs the reference checkout to build oracle/_ref/libref_shaders.so")

PASS_NAMES = ["velocity", "curl", "vorticity", "boundary", "advection", "pressure", "lighting", "precipitation", "iterNum++"]
SAVES = sorted(glob.glob(os.path.join(REFERENCE_SAVES, "*.weathersandbox")), key=os.path.getsize)


def make_ref(g, base, water, wall, drops, fi=None, snd=None):
    h, w = base.shape[:2]
    ref = R.RefShaderSim(w, h, 0 if drops is None else drops.shape[0])
    ref.upload(base, water, wall, drops)
    ref.set_params(P.derive_params(g))
    ref.set_frame_inputs(fi if fi is not None else P.frame_inputs(g))
    ref.set_profiles(P.initial_T_profile(h, g), *(snd or ()))
    return ref


def differences(ora, ref, sun_tolerance=False):
    """Names of the buffers that are not bit-identical (NaN == NaN) between the oracle and the reference shaders."""
    bad = []

    def same(a, b):
        return np.array_equal(a.view(np.uint8), b.view(np.uint8)) or np.array_equal(a, b, equal_nan=True)

    for f, name in ((O.FIELD_BASE, "base"), (O.FIELD_WATER, "water"), (O.FIELD_WALL, "wall"), (O.FIELD_LIGHT, "light")):
        for b in (0, 1):
            a, r = ora.field(f, b, copy=False), ref.field(f, b, copy=False)
            if same(a, r):
                continue
            if name == "light" and sun_tolerance and same(a[..., 1:], r[..., 1:]):
                # the LINEAR fetch's sample position, rounded in normalised vs pixel coordinates (module docstring)
                if np.abs(a[..., 0] - r[..., 0]).max() <= 1e-4 * max(np.abs(r[..., 0]).max(), 1.0):
                    continue
            d = a != r
            bad.append(f"{name}_{b}: {int(d.sum())} values, channels {[int(d[..., c].sum()) for c in range(4)]}, max |d| {np.abs(a.astype(np.float64) - r).max():.3g}")
    for f, name in ((O.FIELD_FEEDBACK, "feedback"), (O.FIELD_DEPOSITION, "deposition"), (O.FIELD_CURL, "curl"), (O.FIELD_VORT, "vortForce")):
        a, r = ora.field(f, 0, copy=False), ref.field(f, 0, copy=False)
        if not same(a, r):
            bad.append(f"{name}: {int((a != r).sum())} values")
    if ora.ND:
        for b in (0, 1):
            a, r = ora.droplets(b, copy=False), ref.droplets(b, copy=False)
            if not same(a, r):
                bad.append(f"droplets_{b}: {int((a != r).any(axis=1).sum())} droplets")
    if not same(ora.lightning, ref.lightning):
        bad.append(f"lightning: {ora.lightning} vs {ref.lightning}")
    if ora.inactive_droplets != ref.inactive_droplets:
        bad.append(f"inactiveDroplets: {ora.inactive_droplets} vs {ref.inactive_droplets}")
    if ora.even != ref.even or ora.iter != ref.iter:
        bad.append("loop state (even / iterNum)")
    return bad


def crop_pow2(sf, x0, w, h):
    """A w x h window (powers of two) of a shipped save, rows from the ground up; droplets re-seeded for the window."""
    cols = np.arange(x0, x0 + w) % sf.width
    base, water, wall = (np.ascontiguousarray(a[:h, cols]) for a in (sf.base, sf.water, sf.wall))
    drops = wsb200.synth.init_rain_drops(wsb200.savefile.num_droplets(w, h), 11)
    return base, water, wall, drops


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", SAVES, ids=[os.path.basename(p)[:-len(".weathersandbox")] for p in SAVES])
def test_every_pass_on_every_shipped_save(path):
    """Each pass of two iterations, with the reference shaders started from the oracle's complete state before
    every pass (second iteration: before its first pass): identical inputs, outputs compared bit for bit (SUNLIGHT:
    module docstring)."""
    sf = wsb200.savefile.load(path)
    g = P.resolve_settings(sf.settings_json)
    ora = make_oracle(g, sf.base, sf.water, sf.wall, sf.droplets)
    ref = make_ref(g, sf.base, sf.water, sf.wall, sf.droplets)
    pow2 = (sf.width & (sf.width - 1)) == 0 and (sf.height & (sf.height - 1)) == 0
    for it in range(2):
        for p, name in enumerate(PASS_NAMES):
            # iteration 0: identical inputs before EVERY pass; iteration 1: before the first only (the one tolerated
            # difference, SUNLIGHT, appears in the lighting pass and nothing after it in the iteration reads the light)
            if it == 0 or p == 0:
                ref.copy_state_from(ora)
            ora.run_pass(p)
            ref.run_pass(p)
            bad = differences(ora, ref, sun_tolerance=not pow2)
            assert not bad, f"{os.path.basename(path)}: iteration {it}, {name} pass: {bad}"


CROPS = [("Hotlake Valley", 900, 512, 128, 300), ("Powerful Hail and Snow Cells", 2000, 512, 128, 120),
         ("Two nice cells in lake valley", 1500, 256, 256, 120), ("Mountain Snow Storm", 300, 1024, 64, 60)]


@pytest.mark.parametrize("name,x0,w,h,iters", CROPS, ids=[c[0] for c in CROPS])
def test_power_of_two_crops_run_bit_identical(name, x0, w, h, iters):
    """Independent runs (no re-synchronisation) from a window of a shipped save, precipitation on: every buffer
    of the loop, the light included, stays bit-identical for the whole run."""
    path = os.path.join(REFERENCE_SAVES, name + ".weathersandbox")
    if not os.path.exists(path):
        pytest.skip("save not in this checkout")
    sf = wsb200.savefile.load(path)
    g = P.resolve_settings(sf.settings_json)
    g["enablePrecipitation"] = True
    base, water, wall, drops = crop_pow2(sf, x0, w, h)
    ora = make_oracle(g, base, water, wall, drops)
    ref = make_ref(g, base, water, wall, drops)
    done = 0
    for n in sorted({1, 2, 10, iters // 2, iters}):
        ora.step(n - done)
        ref.step(n - done)
        done = n
        bad = differences(ora, ref)
        assert not bad, f"{name} {w}x{h}: after {n} iterations: {bad}"
    assert np.isfinite(ora.field(O.FIELD_BASE, 0)).all()
    assert (ora.light_latest()[..., 0] > 0).any()  # the sun has been switched on and travelled


# every tool of advectionShader.frag:229-457: type, intensity, (x, y, size), move, wrap, airplane
TOOL_CASES = [
    (1, 0.7, (0.31, 0.12, 9.0), (0, 0), 0, (0, 0, 0, 0)), (1, -0.4, (-1.0, 0.03, 3.0), (0, 0), 0, (0, 0, 0, 0)),
    (2, 0.25, (0.62, 0.55, 12.0), (0, 0), 1, (0, 0, 0, 0)), (2, -3.0, (0.98, 0.5, 10.0), (0, 0), 1, (0, 0, 0, 0)),
    (3, 1.5, (0.4, 0.3, 8.0), (0, 0), 0, (0.4, 0.32, 0.5, -1.0)), (4, 0.8, (0.5, 0.4, 14.0), (0.02, -0.01), 0, (0, 0, 0, 0)),
    (4, 0.8, (-1.0, 0.4, 5.0), (0.02, -0.01), 0, (0, 0, 0, 0)), (10, 1.0, (0.2, 0.5, 6.0), (0, 0), 0, (0, 0, 0, 0)),
    (11, 1.0, (0.45, 0.35, 7.0), (0, 0), 0, (0, 0, 0, 0)), (12, 1.0, (0.7, 0.2, 7.0), (0, 0), 0, (0, 0, 0, 0)),
    (11, -1.0, (0.3, 0.06, 9.0), (0, 0), 0, (0, 0, 0, 0)),
] + [(t, s, (-1.0, 0.10, 12.0), (0, 0), 0, (0, 0, 0, 0)) for t in (13, 14, 15, 16, 20, 21, 22) for s in (1.0, -1.0)] + [
    (0, 0.0, (0.5, 0.5, 5.0), (0, 0), 1, (0.52, 0.3, 1.0, 1.0)), (0, 0.0, (0.5, 0.5, 5.0), (0, 0), 0, (0.30, 0.10, 1.0, 1.0)),
]


def _stress_pair(seed, w=128, h=64):
    g, base, water, wall, drops = stress_state(w, h, seed=seed)
    g["enablePrecipitation"] = True
    g["soundingForcing"] = 0.95
    g["globalDrying"] = 0.00002
    g["globalHeating"] = 0.0001
    sea = wall[0, :, 0] == 2
    for x0, t in ((0, 5), (16, 6), (32, 4), (48, 3), (64, 1)):  # runway, industrial, urban, fire, land
        cols = np.arange(x0, x0 + 16)
        wall[:, cols[~sea[cols]], 0] = t
    rng = np.random.default_rng(2)
    snd = (P.initial_T_profile(h, g) + rng.normal(0, 1.0, h + 1).astype(f32), rng.uniform(0, 8, h + 1).astype(f32), rng.normal(0, 0.05, h + 1).astype(f32))
    ora = make_oracle(g, base, water, wall, drops)
    ora.set_profiles(P.initial_T_profile(h, g), *snd)
    ref = make_ref(g, base, water, wall, drops, snd=snd)
    return g, ora, ref


@pytest.mark.parametrize("case", TOOL_CASES, ids=[f"type{c[0]}{'+' if c[1] >= 0 else '-'}{i}" for i, c in enumerate(TOOL_CASES)])
def test_every_tool_and_the_airplane_bit_identical(case):
    """A stress state (every wall type, fire, snow, desert, smoke, clouds, droplets) on a 128 x 64 grid with sounding
    forcing, global drying and heating switched on: three idle iterations, three with the tool / airplane input,
    two idle — every buffer bit-identical."""
    kind, intensity, (bx, by, size), move, wrap, plane = case
    g, ora, ref = _stress_pair(29)
    fi = P.frame_inputs(g)
    fi.userInputType = kind
    for k, v in enumerate((bx, by, intensity, size)):
        fi.userInputValues[k] = v
    fi.userInputMove[0], fi.userInputMove[1] = move
    fi.wrapHorizontally = wrap
    for k, v in enumerate(plane):
        fi.airplaneValues[k] = v
    compared = 0
    for inputs, n in ((P.frame_inputs(g), 3), (fi, 3), (P.frame_inputs(g), 2)):
        ora.set_frame_inputs(inputs)
        ref.set_frame_inputs(inputs)
        for _ in range(n):
            ora.step(1)
            ref.step(1)
            if not all(np.isfinite(ora.field(f, b, copy=False)).all() for f in (O.FIELD_BASE, O.FIELD_WATER) for b in (0, 1)):
                # the whole-width "remove" tools 13-16 leave burning columns that blow the reference's own arithmetic up
                # to NaN within two iterations; parity is claimed for finite states
                assert compared >= 4, "the state left the finite range before the tool was applied"
                return
            bad = differences(ora, ref)
            assert not bad, f"type {kind}, intensity {intensity}, iteration {ora.iter}: {bad}"
            compared += 1


def test_slow_processes_at_iteration_multiples_bit_identical():
    """The same stress state run across iterations 9990 .. 10110: growth ticks, snow / soil-moisture smoothing, fire
    spread and burn-out (multiples of 100 and of 10000, boundaryShader.frag:409-472), dynamic water temperature
    (every 20), with forcing on and precipitation running."""
    g, ora, ref = _stress_pair(31)
    p = P.derive_params(g)
    p.dynamicWaterTemperature = 1.0
    ora.set_params(p)
    ref.set_params(p)
    ora.iter = ref.iter = 9990
    for _ in range(12):
        ora.step(10)
        ref.step(10)
        bad = differences(ora, ref)
        assert not bad, f"iteration {ora.iter}: {bad}"
    assert np.isfinite(ora.field(O.FIELD_BASE, 0)).all()


def test_precipitation_life_cycle_bit_identical():
    """Dense cold cloud aloft and warm cloud near the ground on a 64 x 64 grid: droplets spawn as rain and snow,
    grow, freeze, melt, fall and deposit, across iteration 600 (the inactive-count
    latch) — droplet records, feedback / deposition textures and both latches bit-identical throughout."""
    w, h = 64, 64
    g, base, water, wall, _ = stress_state(w, h, seed=8)
    g["enablePrecipitation"] = True
    rng = np.random.default_rng(8)
    air = wall[..., 1] != 0
    water[h // 2:, :, 1] += np.where(air[h // 2:], f32(30.0), f32(0.0))
    water[h // 2:, :, 0] += np.where(air[h // 2:], f32(30.0), f32(0.0))
    water[:h // 5, :, 1] += np.where(air[:h // 5], f32(40.0), f32(0.0))
    water[:h // 5, :, 0] += np.where(air[:h // 5], f32(40.0), f32(0.0))
    n = 500
    drops = np.zeros((n, 5), f32)
    drops[:, 0] = rng.uniform(-1, 1, n)
    drops[:, 1] = rng.uniform(-0.95, 0.95, n)
    drops[:, 2] = rng.uniform(0.0, 0.5, n)
    drops[:, 3] = rng.uniform(0.0, 0.5, n)
    drops[:, 4] = rng.choice([0.2, 0.6, 1.0], n)
    drops[:200, 2] = -2.0 - rng.uniform(-1, 1, 200)         # inactive, position kept as seed
    drops[200:240, 2:4] = rng.uniform(0.0, 0.015, (40, 2))  # residual droplets: evaporate
    drops[240:280, 1] = rng.uniform(-1.0, -0.93, 40)        # in / just above the ground
    drops[280:290, 1] = -1.001                              # below the map
    drops[290:340, 2] = 0.0                                 # pure ice
    drops[340:380, 3] = 0.0                                 # pure rain
    p = P.derive_params(g)
    p.spawnChanceMult = 0.02
    ora = make_oracle(g, base, water, wall, drops)
    ref = make_ref(g, base, water, wall, drops)
    ora.set_params(p)
    ref.set_params(p)
    ora.iter = ref.iter = 597
    for _ in range(12):
        ora.step(5)
        ref.step(5)
        bad = differences(ora, ref)
        assert not bad, f"iteration {ora.iter}: {bad}"
    assert (ora.droplets()[:, 2] >= 0).sum() != (drops[:, 2] >= 0).sum()  # droplets were spawned / retired


@pytest.mark.parametrize("seed,cold_cloud", [(26, 7.0), (26, 5.0), (25, 5.0), (22, 5.0)])
def test_lightning_bolt_bit_identical(seed, cold_cloud):
    """Sub-zero cloud aloft and a pool of inactive droplets: a spawn turns into a lightning bolt
    (precipitationShader.vert:121-140: 1-pixel sprite into feedback pixel (1, 0), START_ITERNUM = iterNum - initalMass
    through the shared VAPOR channel) and lightningLocationShader.frag:24-38 latches it, or discards."""
    w, h = 64, 64
    g, base, water, wall, _ = stress_state(w, h, seed=seed)
    g["enablePrecipitation"] = True
    rng = np.random.default_rng(seed)
    air = wall[..., 1] != 0
    water[h // 2:, :, 1] += np.where(air[h // 2:], f32(cold_cloud), f32(0.0))
    water[h // 2:, :, 0] += np.where(air[h // 2:], f32(cold_cloud), f32(0.0))
    n = 400
    drops = np.zeros((n, 5), f32)
    drops[:, 0] = rng.uniform(-1, 1, n)
    drops[:, 1] = rng.uniform(-0.95, 0.95, n)
    drops[:, 2] = -2.0 - rng.uniform(-1, 1, n)
    drops[:, 4] = 1.0
    p = P.derive_params(g)
    p.spawnChanceMult = 0.02
    ora = make_oracle(g, base, water, wall, drops)
    ref = make_ref(g, base, water, wall, drops)
    ora.set_params(p)
    ref.set_params(p)
    ora.iter = ref.iter = 599  # the second iteration is a multiple of 600: the inactive count is latched (app.js:5957-5967)
    struck = 0
    for _ in range(40):  # long enough for a second bolt (30 iterations must pass, :131)
        ora.step(1)
        ref.step(1)
        bad = differences(ora, ref)
        assert not bad, f"iteration {ora.iter}: {bad}"
        struck = max(struck, ref.lightning[2])
    assert struck > 598.0 and ref.lightning[3] > 0.0, "no lightning bolt was latched: the branch was not exercised"
    assert ref.inactive_droplets > 0


def test_growth_rates_beyond_100_take_the_canonical_modulo():
    """Soil moisture 250 .. 900 under a high sun: vegetationGrowthRate > 100, growth interval (100 / rate) * 100 == 0,
    `%` by zero (boundaryShader.frag:460).  GLSL leaves it undefined; frozen as "no growth tick" — here the reference's
    own expression runs with that one definition (glsl_shim.h glsl_nz) and agrees with the oracle at iterations 0, 100
    and 10000."""
    w, h = 128, 32
    g, base, water, wall, _ = stress_state(w, h, seed=17)
    g["sunAngle"] = 85.0
    g["enablePrecipitation"] = False
    land = (wall[..., 1] == 0) & np.isin(wall[..., 0], (1, 3, 4, 6))
    water[..., 2] = np.where(land, f32(250.0) + f32(650.0) * np.random.default_rng(5).random((h, w)).astype(f32), water[..., 2])
    ora = make_oracle(g, base, water, wall, None)
    ref = make_ref(g, base, water, wall, None)
    for start in (0, 99, 9999):
        ora.iter = ref.iter = start
        for b in (0, 1):  # sunlight travels one row per iteration: put it at the surface
            ora.field(O.FIELD_LIGHT, b, copy=False)[..., 0] = f32(900.0)
            ref.field(O.FIELD_LIGHT, b, copy=False)[..., 0] = f32(900.0)
        ora.step(2)
        ref.step(2)
        bad = differences(ora, ref)
        assert not bad, f"from iteration {start}: {bad}"


def test_grid_taller_than_the_reference_cap_bit_identical():
    """512 rows: beyond the 503 the reference's own 126-vec4 profile tables allow.  The translator raises
    ONLY that array bound (translate.py); the shaders index it as they always did.  A power-of-two stress state, 40
    iterations with particles: every buffer bit-identical to the oracle, whose tables are sized by the grid."""
    w, h = 128, 512
    g, base, water, wall, drops = stress_state(w, h, seed=5)
    g["enablePrecipitation"] = True
    ora = make_oracle(g, base, water, wall, drops)
    ref = make_ref(g, base, water, wall, drops)
    for n in (1, 9, 30):
        ora.step(n)
        ref.step(n)
        bad = differences(ora, ref)
        assert not bad, f"iteration {ora.iter}: {bad}"
    assert np.isfinite(ora.field(O.FIELD_BASE, 0)).all()
    with pytest.raises(ValueError):
        R.RefShaderSim(16, 5000)


def test_full_width_grid_coordinate_quantisation_bit_identical():
    """16384 columns (the BASELINE width) x 64 rows: beyond x = 8192 the fp32 spacing of `fragCoord - v` is 2^-10 of a
    cell and the back-trace of advectionShader.frag:93-103 / common.glsl:194-254 quantises accordingly (hard
    part 2) — the oracle, whose large-grid strips are the GPU tests' bar at 16384 x 4096, against the shader code itself."""
    w, h = 16384, 64
    g = P.resolve_settings(None)
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    g["enablePrecipitation"] = False
    base, water, wall, _ = wsb200.synth.full_state(w, h, seed=7, g=g, with_droplets=False, vel_amplitude=0.3)
    ora = make_oracle(g, base, water, wall, None)
    ref = make_ref(g, base, water, wall, None)
    for n in (1, 5):
        ora.step(n)
        ref.step(n)
        bad = differences(ora, ref)
        assert not bad, f"iteration {ora.iter}: {bad}"
    moved = ora.field(O.FIELD_BASE, 0)[:, 8192:, :2]
    assert np.abs(moved).max() > 0.05  # there is flow to trace back in the quantised half


@pytest.mark.parametrize("scale", [8.0, 40.0])
def test_fast_flow_back_trace_bit_identical(scale):
    """Velocities of 0.45 and 2.3 cells per iteration (the shipped saves peak at 0.36): the back-trace leaves the
    neighbouring cell, bilerp / bilerpWall fetch texels several cells away and across the periodic seam."""
    w, h = 256, 64
    base, water, wall = wsb200.synth.dry_state(w, h, seed=1234)
    base[..., :2] *= f32(scale)
    wall[20:30, 60:70, 1] = 0  # a block of wall in the flow: the wall-aware weights of bilerpWall
    wall[20:30, 60:70, 0] = 1
    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    ora = make_oracle(g, base, water, wall, None)
    ref = make_ref(g, base, water, wall, None)
    assert np.abs(base[..., :2]).max() > 0.05 * scale
    for n in (1, 3):
        ora.step(n)
        ref.step(n)
        bad = differences(ora, ref)
        assert not bad, f"scale {scale}, iteration {ora.iter}: {bad}"


@pytest.mark.parametrize("w,h", [(128, 64), (2048, 32), (64, 1024)])
def test_sun_fetch_bit_identical_at_every_sun_angle_on_power_of_two_grids(w, h):
    """The LINEAR fetch `texture(lightTex, texCoord + sunRay)` with the sun at the zenith (ray offset exactly one row),
    2 degrees off it (an offset a few ulp short of a texel centre at large y), near and below the horizon, from either
    side: normalised (shader) and pixel-space (frozen) arithmetic give the same bits when 1 / size is a power of two."""
    for gui_angle in (90.0, 88.0, 60.0, 1.0, -3.0, 172.0):
        g, base, water, wall, _ = stress_state(w, h, seed=3)
        g["sunAngle"] = gui_angle
        g["enablePrecipitation"] = False
        ora = make_oracle(g, base, water, wall, None)
        ref = make_ref(g, base, water, wall, None)
        ora.step(24)
        ref.step(24)
        bad = differences(ora, ref)
        assert not bad, f"{w}x{h}, sun at {gui_angle} degrees: {bad}"


def test_dry_sweep_equals_the_reference_passes_it_is_made_of():
    """BASELINE config 2 (the >= 70 % roofline kernel k_fused_dry) is velocity -> advection -> pressure of the BASE field
    with the boundary pass left out — a schedule the bench defines, made of reference passes.  oracle.step_dry (the
    GPU dry sweep's bar) against the reference's velocityShader / advectionShader / pressureShader run in that order."""
    w, h, iters = 256, 64, 12
    base, water, wall = wsb200.synth.dry_state(w, h, seed=1234)
    base[..., :2] *= f32(3.0)
    wall[24:34, 100:120, 1] = 0
    wall[24:34, 100:120, 0] = 1
    g = P.resolve_settings(None)
    g["dragMultiplier"], g["wind"] = 0.001, 0.0
    g["enablePrecipitation"] = False
    ora = make_oracle(g, base, water, wall, None)
    ref = make_ref(g, base, water, wall, None)
    for _ in range(iters):
        ref.run_pass(0)  # velocityShader: frameBuff_0 -> frameBuff_1
        ref.field(O.FIELD_BASE, 0, copy=False)[...] = ref.field(O.FIELD_BASE, 1, copy=False)  # no boundary pass in between
        ref.field(O.FIELD_WALL, 0, copy=False)[...] = ref.field(O.FIELD_WALL, 1, copy=False)
        ref.run_pass(4)  # advectionShader: frameBuff_0 -> frameBuff_1
        ref.run_pass(5)  # pressureShader:  frameBuff_1 -> frameBuff_0
        ref.run_pass(8)  # iterNum++
    ora.step_dry(iters)
    got, want = ora.field(O.FIELD_BASE, 0), ref.field(O.FIELD_BASE, 0)
    assert np.array_equal(got, want), f"dry sweep: base differs in {(got != want).sum()} values"
    assert np.abs(want[..., :2]).max() > 0.05


@pytest.mark.parametrize("w,h,seed,mult", [(256, 128, 0.37, 0.5), (300, 100, 0.81, 0.9), (128, 64, 0.5, 0.07), (64, 64, 0.1, 0.01), (1000, 250, 0.2566, 0.33), (2000, 300, 0.37, 1.0)])
def test_setup_shader_matches_synth_setup_state(w, h, seed, mult):
    """setupShader.frag (the reference's own, compiled) drawn once == synth.setup_state, bit for bit."""
    g = P.resolve_settings(None)
    base, water, wall, _ = wsb200.synth.setup_state(w, h, seed=seed, height_mult=mult, g=g, with_droplets=False)
    rb, rw, rwl = R.setup_state(w, h, seed, mult, float(g["simHeight"]), float(P.dry_lapse(g)), P.initial_T_profile(h, g))
    assert np.array_equal(wall, rwl), f"wall differs in {(wall != rwl).sum()} bytes"
    assert np.array_equal(base, rb), f"base differs in {(base != rb).sum()} values"
    assert np.array_equal(water, rw), f"water differs in {(water != rw).sum()} values"
