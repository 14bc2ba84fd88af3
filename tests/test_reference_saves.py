"""Pins of the oracle against REFERENCE-PRODUCED data: the shipped saves.

Every `saves/*.weathersandbox` of the reference checkout is the reference's own WebGL output
(prepareDownload, app.js:6575-6628, reads frameBuff_0 and droplet buffer 0 back from the GPU): the
state after an unknown number of iterations of its shaders.  Nothing else the reference produced is
available (no tests, no golden vectors, and it cannot run in this image), so these files
are what the oracle — and through it the CUDA path — can be held to:

  * the integer wall planes TYPE, DISTANCE and VERT_DISTANCE of a save are a FIXED POINT of one more
    iteration (the distance propagation, saturation and dyke rules of boundaryShader.frag:92-269,
    the below-surface counters :372-390, the pass-throughs of advection / pressure): 0 changed bytes,
    except LAND <-> FIRE flips in the saves that are burning (fire spread / burn-out, :403-413, :469-472);
  * VEGETATION of every land / urban / fire SURFACE cell (the stored state) is unchanged by an
    iteration that is not a growth tick; the cells that do change are exactly the derived copies —
    air cells 1..8 above the ground ("copied from below", :310), cells below the surface ("copied
    from above", :379) and WATER surface cells set to 20 (:524) — which the saves hold in the form an
    OLDER revision of the shaders wrote (the same revision that marked wall cells with water = 1111,
    today commented out at advectionShader.frag:199); they settle within the 8-row copy chain;
  * reference-written invariants an iteration must keep: zero velocity inside walls, wall marker <->
    wall mask, the T == 1000 snow-melt convention of LAND wall cells (advectionShader.frag:195-197),
    water-surface temperatures inside the clamp of boundaryShader.frag:522, finite fields, |v| < 1;
  * continuity: a save is a snapshot of a running simulation, so one more iteration of a faithful
    restatement moves the fields by a small fraction of their spread (a wrong sign or coefficient in
    the pressure / velocity / advection chain shows up as a jump).

On a box without /root/reference (the GPU box) the same checks run on the files committed under
tests/golden/: the 100 x 100 save and two 512-column crops (make_save_crops.py) — on the crops away
from the artificial periodic seam."""
import glob
import os

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

P = wsb200.params
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REFERENCE_SAVES = os.environ.get("WSB_REFERENCE_SAVES", "/root/reference/saves")  # the GPU box has no reference checkout
SEAM_MARGIN = 136  # crops: DISTANCE saturates at 127, so the seam cannot reach further
TYPE, DISTANCE, VERT_DISTANCE, VEGETATION = range(4)
LAND, WATER, FIRE = 1, 2, 3


def available_saves():
    """(id, path, seam margin): every reference save when the checkout is present, plus the committed fixtures."""
    out = []
    for p in sorted(glob.glob(os.path.join(REFERENCE_SAVES, "*.weathersandbox")), key=os.path.getsize):
        out.append((os.path.basename(p)[:-len(".weathersandbox")], p, 0))
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.weathersandbox"))):
        name = os.path.basename(p)[:-len(".weathersandbox")]
        out.append(("golden/" + name, p, SEAM_MARGIN if "crop" in name else 0))
    return out


SAVES = available_saves()


def oracle_from_save(sf, iter_num=0):
    g = P.resolve_settings(sf.settings_json)
    ora = O.OracleSim(sf.width, sf.height, sf.droplets.shape[0])
    ora.upload(sf.base, sf.water, sf.wall, sf.droplets)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(sf.height, g))
    ora.iter = iter_num
    return ora, g


def interior(a, margin):
    return a if margin == 0 else a[:, margin:a.shape[1] - margin]


def check_wall_pins(sf, wall_after, margin, what):
    """The fixed-point pins on the wall texel after ONE iteration (iterNum not a growth tick)."""
    w0, w1 = interior(sf.wall, margin), interior(wall_after, margin)
    assert np.array_equal(w1[..., DISTANCE], w0[..., DISTANCE]), f"{what}: DISTANCE is not a fixed point ({(w1[..., 1] != w0[..., 1]).sum()} cells)"
    assert np.array_equal(w1[..., VERT_DISTANCE], w0[..., VERT_DISTANCE]), f"{what}: VERT_DISTANCE is not a fixed point"
    dt = w1[..., TYPE] != w0[..., TYPE]
    flips = set(zip(w0[..., TYPE][dt].tolist(), w1[..., TYPE][dt].tolist()))
    assert flips <= {(LAND, FIRE), (FIRE, LAND)}, f"{what}: wall TYPE changed other than by fire: {sorted(flips)}"
    if not (w0[..., TYPE] == FIRE).any():
        assert not dt.any(), f"{what}: TYPE changed in {dt.sum()} cells of a save without fire"
    # vegetation: the stored state (surface cells of land kinds) stays; only the derived copies move
    is_wall = w0[..., DISTANCE] == 0
    above_air = np.roll(sf.wall[..., DISTANCE], -1, axis=0) != 0
    above_air = interior(above_air, margin)
    dv = w1[..., VEGETATION] != w0[..., VEGETATION]
    surface_land = is_wall & above_air & (w0[..., TYPE] != WATER) & (w0[..., TYPE] != FIRE) & ~dt
    assert not (dv & surface_land).any(), f"{what}: vegetation of {(dv & surface_land).sum()} land surface cells changed outside a growth tick"
    water_surface = dv & is_wall & above_air & (w0[..., TYPE] == WATER)
    assert (w1[..., VEGETATION][water_surface] == 20).all()                      # boundaryShader.frag:524
    air = dv & ~is_wall
    assert ((w0[..., VERT_DISTANCE][air] >= 1) & (w0[..., VERT_DISTANCE][air] <= 8)).all(), f"{what}: vegetation changed in air above the copy zone"
    return int(dv.sum())


@pytest.mark.parametrize("name,path,margin", SAVES, ids=[s[0] for s in SAVES])
def test_save_is_a_fixed_point_of_the_wall_planes(name, path, margin):
    sf = wsb200.savefile.load(path)
    ora, _ = oracle_from_save(sf, iter_num=7)  # not a multiple of the slow processes' intervals
    ora.step(1)
    w1 = ora.field(O.FIELD_WALL, 0)
    check_wall_pins(sf, w1, margin, name)
    # the derived vegetation copies settle within the 8-row chain (+ the surface row, + 1 to see it); burning saves keep burning
    burning = (sf.wall[..., TYPE] == FIRE).any()
    prev = w1
    for n in range(2, 13):
        ora.step(1)
        cur = ora.field(O.FIELD_WALL, 0)
        changed = int((interior(cur, margin)[..., VEGETATION] != interior(prev, margin)[..., VEGETATION]).sum())
        prev = cur
        if changed == 0:
            break
    assert burning or (changed == 0 and n <= 10), f"{name}: vegetation still changing after {n} iterations ({changed} cells)"
    assert np.array_equal(interior(cur, margin)[..., 1:3], interior(sf.wall, margin)[..., 1:3]), f"{name}: distances moved within {n} iterations"


@pytest.mark.parametrize("name,path,margin", SAVES, ids=[s[0] for s in SAVES])
def test_reference_written_invariants_survive_an_iteration(name, path, margin):
    sf = wsb200.savefile.load(path)
    wall_cells = sf.wall[..., DISTANCE] == 0
    # what the reference wrote
    assert (sf.wall[0, :, DISTANCE] == 0).all()                                  # row 0 is always wall
    assert (sf.base[..., 0:2][wall_cells] == 0).all()                            # no velocity inside walls
    marker = np.unique(sf.water[..., 0][wall_cells])
    assert set(marker.tolist()) <= {1111.0, 1001.0, 1002.0}                      # wall marker (1111: older shader revision)
    assert not np.isin(sf.water[..., 0][~wall_cells], (1111.0, 1001.0, 1002.0)).any()   # marker <=> wall mask
    land_T = sf.base[..., 3][wall_cells & (sf.wall[..., TYPE] == LAND)]
    assert np.all(np.abs(land_T - 1000.0) < 1e-2)                                # snow-melt convention
    surface_water = wall_cells & (sf.wall[..., TYPE] == WATER) & (np.roll(sf.wall[..., DISTANCE], -1, axis=0) != 0)
    sw_T = sf.base[..., 3][surface_water]
    assert sw_T.size == 0 or (sw_T.min() >= 273.15 - 1e-3 and sw_T.max() <= 273.15 + 40.0 + 1e-3)   # boundaryShader.frag:522
    vmax0 = float(np.abs(sf.base[..., 0:2]).max())
    assert vmax0 < 1.0
    # ... and what one more iteration (oracle) makes of it
    ora, _ = oracle_from_save(sf, iter_num=7)
    ora.step(1)
    base, water, wall = ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 0)
    assert np.isfinite(base).all() and np.isfinite(water).all() and np.isfinite(ora.light_latest()).all()
    wc = wall[..., DISTANCE] == 0
    assert np.array_equal(wc, wall_cells)
    assert (base[..., 0:2][wc] == 0).all()
    assert (water[..., 0][wc] == np.where(wall[..., TYPE][wc] == WATER, np.float32(1002.0), np.float32(1001.0))).all()   # advectionShader.frag:403-409
    # LAND wall cells carry T = 1000 (advectionShader.frag:195-197); the SURFACE cell subtracts the snow-melt cooling the
    # pressure pass hands to the air above (advectionShader.frag:214-228, pressureShader.frag:24-27), so only buried cells are exact
    buried = wc & (wall[..., TYPE] == LAND) & (np.roll(wall[..., DISTANCE], -1, axis=0) == 0)
    assert (base[..., 3][buried] == 1000.0).all()
    assert np.all(np.abs(base[..., 3][wc & (wall[..., TYPE] == LAND)] - 1000.0) < 1e-2)
    sw1 = base[..., 3][surface_water]
    assert sw1.size == 0 or (sw1.min() >= np.float32(273.15) and sw1.max() <= np.float32(273.15 + 40.0))
    assert (water[..., 0][~wc] >= 0).all()
    # continuity of a running simulation: one iteration moves each field by a small part of its spread in the air
    # (measured on the 13 weather saves: vx 0.2-1.2 %, vy 0.7-12 %, T 0.007-0.14 % of the field's spread; the 100 x 100
    # test save is nearly at rest and the crops carry their seam, so they only get the loose bound)
    air = interior(~wc, margin)
    full = margin == 0 and sf.width >= 1000
    for ch, cname, bound in ((0, "vx", 0.03 if full else 0.5), (1, "vy", 0.2 if full else 0.5), (3, "T", 0.005 if full else 0.01)):
        a0, a1 = interior(sf.base[..., ch], margin)[air], interior(base[..., ch], margin)[air]
        spread = float(np.std(a0)) + 1e-6
        rms = float(np.sqrt(np.mean((a1.astype(np.float64) - a0) ** 2)))
        assert rms < bound * spread + 1e-4, f"{name}: {cname} moved by rms {rms:.4g} in one iteration (spread {spread:.4g})"
    tot0 = float(interior(sf.water[..., 0], margin)[air].sum(dtype=np.float64))
    tot1 = float(interior(water[..., 0], margin)[air].sum(dtype=np.float64))
    assert abs(tot1 - tot0) < 2e-4 * tot0, f"{name}: total water in the air jumped by {(tot1 - tot0) / tot0:.3g} in one iteration"


def test_growth_tick_explains_the_iteration_zero_residual():
    """Loaded saves restart at iterNum = 0 (app.js:440), which IS a growth tick for every interval: there the land
    surface vegetation may move by +1 (boundaryShader.frag:460-465) and nowhere by more."""
    _, path, margin = SAVES[0] if len(SAVES) == 1 else [s for s in SAVES if "100" in s[0]][0]
    sf = wsb200.savefile.load(path)
    ora, _ = oracle_from_save(sf, iter_num=0)
    ora.step(1)
    w1 = ora.field(O.FIELD_WALL, 0)
    is_wall = sf.wall[..., DISTANCE] == 0
    surface_land = is_wall & (np.roll(sf.wall[..., DISTANCE], -1, axis=0) != 0) & (sf.wall[..., TYPE] == LAND)
    d = w1[..., VEGETATION].astype(int) - sf.wall[..., VEGETATION]
    assert set(np.unique(d[surface_land]).tolist()) <= {0, 1}
