"""Independent restatement of the precipitation-particle pass (precipitationShader.vert:66-298,
precipitationShader.frag, app.js:5933-5952) as a scalar, droplet-by-droplet Python transliteration
of the GLSL: random respawn sampling (common.glsl:103-137 hash / random2d), spawn chance, lightning
spawn and latch (lightningLocationShader.frag), the inactive-droplet count latch, growth / freezing / melting / evaporation with their feedback to the fluid, fall and
horizontal wrap, deposition on the ground; then the frozen rasterisation rule of DESIGN.md 2
(point sprites cover the pixels whose centre lies in [c - size/2, c + size/2), clipped, never
wrapped, `out` varyings start at zero, blend ONE, ONE in droplet order) and pow(m, 1/3) as the
bit-guess + four Newton steps of DESIGN.md 2.  The C++ oracle must reproduce the droplet records
AND the feedback / deposition textures bit for bit."""
import math

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
u32 = np.uint32
Z, ONE = f32(0.0), f32(1.0)
MASS, HEAT, VAPOR = 0, 1, 2
K0 = f32(0.0) + f32(273.15)  # CtoK(0.0)


def _hash(x):
    x = int(x) & 0xFFFFFFFF
    x = (x + (x << 10)) & 0xFFFFFFFF
    x ^= x >> 6
    x = (x + (x << 3)) & 0xFFFFFFFF
    x ^= x >> 11
    x = (x + (x << 15)) & 0xFFFFFFFF
    return x


def _bits(f):
    return int(np.array([f], f32).view(u32)[0])


def _from_bits(u):
    return np.array([u & 0xFFFFFFFF], u32).view(f32)[0]


def random2d(sx, sy):
    h = _hash((_bits(sx) + _hash(_bits(sy))) & 0xFFFFFFFF)
    r2 = _from_bits((h & 0x007FFFFF) | 0x3F800000)
    return r2 - ONE * np.floor(r2 / ONE)  # mod(r2, 1.0)


def gmax(a, b):
    return b if a < b else a


def gmin(a, b):
    return b if b < a else a


def map_range(v, min1, max1, min2, max2):
    return min2 + (v - min1) * (max2 - min2) / (max1 - min1)


def max_water(T):
    x = T / f32(250.0)
    x2 = x * x
    x4 = x2 * x2
    x8 = x4 * x4
    return (x8 * x8) * x


def cbrt(m):
    if not m > Z:
        return Z
    y = _from_bits(_bits(m) // 3 + 709921077)
    for _ in range(4):
        y = y - (y - m / (y * y)) * (ONE / f32(3.0))
    return y


def precipitation(base1, water1, drops, lightning, p, iter_num, inactive):
    h, w = base1.shape[:2]
    wf, hf = f32(w), f32(h)
    texel_x, texel_y = f32(1.0 / w), f32(1.0 / h)
    it = f32(iter_num)
    out = np.empty_like(drops)
    seen = {"spawn_rain": 0, "spawn_snow": 0, "lightning": 0, "still_inactive": 0, "tiny": 0, "deposited": 0, "moved_up": 0,
            "freezing": 0, "melting": 0}
    fb_tex = np.zeros((h, w, 4), f32)
    dep_tex = np.zeros((h, w, 2), f32)

    def fetch(tex, tx, ty):  # NEAREST, REPEAT
        return tex[int(np.floor(ty * hf)) % h, int(np.floor(tx * wf)) % w]

    def splat(px, py, size, fbv, depv):
        if not (px >= f32(-1.0) and px <= ONE and py >= f32(-1.0) and py <= ONE):
            return
        xw, yw = (px + ONE) * f32(0.5) * wf, (py + ONE) * f32(0.5) * hf
        half = f32(size) * f32(0.5)
        for j in range(h):
            if not (yw - half <= f32(j) + f32(0.5) < yw + half):
                continue
            for i in range(w):
                if xw - half <= f32(i) + f32(0.5) < xw + half:
                    fb_tex[j, i] += fbv
                    dep_tex[j, i] += depv

    for n in range(drops.shape[0]):
        drop_x, drop_y, mass_w, mass_i, density = (f32(v) for v in drops[n])
        new_x, new_y, new_w, new_i, new_d = drop_x, drop_y, mass_w, mass_i, density
        feedback = [Z, Z, Z, Z]
        deposition = [Z, Z]
        active, spawned, lightning_spawned = True, False, False
        gl_x, gl_y, size = Z, Z, 1
        tx = ty = Z
        base = water = None
        real_temp = Z

        def disable():
            return f32(-2.0) - drop_x, drop_y

        if mass_w < Z:  # inactive: try to respawn
            tx = random2d(mass_w, drop_x + it * f32(0.3754))
            ty = random2d(mass_i, drop_x + it * f32(0.073162))
            base = [f32(v) for v in fetch(base1, tx, ty)]
            water = [f32(v) for v in fetch(water1, tx, ty)]
            real_temp = base[3] - ty * f32(p.dryLapse)
            threshold = f32(p.aboveZeroThreshold) if real_temp > K0 else f32(p.subZeroThreshold)
            if water[1] > threshold and base[3] < f32(500.0):
                chance = ((water[1] - threshold) / (f32(inactive) + f32(10.0))) * wf * hf * f32(p.spawnChanceMult)
                c10 = water[1] * f32(10.0)
                sq = c10 * c10
                nrm = sq - np.floor(sq)
                if chance > nrm:
                    spawned = True
                    new_x, new_y = (tx - f32(0.5)) * f32(2.0), (ty - f32(0.5)) * f32(2.0)
                    seen["spawn_snow" if real_temp < K0 else "spawn_rain"] += 1
                    if real_temp < K0:
                        new_w, new_i = Z, f32(0.15)
                        feedback[HEAT] = feedback[HEAT] + new_i * f32(p.meltingHeat)
                        new_d = f32(p.snowDensity)
                        dens = water[1] + water[2]
                        l_chance = gmax((dens - f32(2.5)) * f32(0.0033), Z)
                        if f32(lightning[2]) < it - f32(30.0) and random2d(base[3] * f32(0.2324), water[0] * f32(7.7)) < l_chance:
                            lightning_spawned = True
                            seen["lightning"] += 1
                            active = False
                            size = 1
                            feedback[0], feedback[1] = tx, ty
                            feedback[2] = it
                            feedback[3] = gmin(gmax(dens / f32(10.0) + (random2d(tx, ty) - f32(0.5)), f32(0.01)), f32(4.0))
                            gl_x, gl_y = f32(-1.0) + texel_x * f32(3.0), f32(-1.0) + texel_y
                    else:
                        new_w, new_i, new_d = f32(0.15), Z, ONE
                    feedback[VAPOR] = feedback[VAPOR] - f32(0.15)
            if spawned:
                if not lightning_spawned:
                    size = 1
                    gl_x, gl_y = new_x, new_y
            else:
                active = False
                seen["still_inactive"] += 1
                size = 1
                feedback[MASS] = ONE
                gl_x, gl_y = f32(-1.0) + texel_x, f32(-1.0) + texel_y

        if active:
            if not spawned:
                tx, ty = drop_x / f32(2.0) + f32(0.5), drop_y / f32(2.0) + f32(0.5)
                water = [f32(v) for v in fetch(water1, tx, ty)]
                base = [f32(v) for v in fetch(base1, tx, ty)]
                real_temp = base[3] - ty * f32(p.dryLapse)
            total = new_w + new_i
            if total < f32(0.04):
                feedback[HEAT] = -(total * f32(p.evapHeat))
                feedback[VAPOR] = total
                seen["tiny"] += 1
                new_w, new_i = disable()
            elif new_y < f32(-1.0) or water[0] > f32(1000.0):
                if f32(fetch(base1, tx, ty + texel_y)[3]) > f32(500.0):
                    new_y = new_y + texel_y * ONE
                    seen["moved_up"] += 1
                seen["deposited"] += 1
                deposition[0], deposition[1] = new_w, new_i
                new_w, new_i = disable()
            else:
                area = cbrt(total)
                rate = gmax(map_range(real_temp, K0, f32(-30.0) + f32(273.15), f32(p.growthRate0C), f32(p.growthRate_30C)), f32(p.growthRate0C))
                growth = water[1] * rate * area
                if real_temp < K0 and water[1] > Z and density == ONE:
                    growth = growth + area * water[2] * f32(0.0030)
                feedback[VAPOR] = feedback[VAPOR] - growth * ONE
                seen["freezing" if real_temp < K0 else "melting"] += 1
                if real_temp < K0:
                    new_i = new_i + growth
                    feedback[HEAT] = feedback[HEAT] + growth * f32(p.meltingHeat)
                    freezing = gmin((K0 - real_temp) * f32(p.freezingRate) * area, new_w)
                    new_w = new_w - freezing
                    new_i = new_i + freezing
                    feedback[HEAT] = feedback[HEAT] + freezing * f32(p.meltingHeat)
                else:
                    new_w = new_w + growth
                    melting = gmin((real_temp - K0) * f32(p.meltingRate) * area, new_i)
                    new_i = new_i - melting
                    new_w = new_w + melting
                    feedback[HEAT] = feedback[HEAT] - melting * f32(p.meltingHeat)
                    new_d = gmin(new_d + (melting / total) * ONE, ONE)
                drop_t = base[3] - ty * f32(p.dryLapse)
                if new_i > Z:
                    drop_t = gmin(drop_t, K0)
                evap_subli = gmax((max_water(drop_t) - water[0]) * area * f32(p.evapRate), Z)
                evap = gmin(new_w, evap_subli)
                subli = gmin(new_i, evap_subli - evap)
                new_w = new_w - evap
                new_i = new_i - subli
                feedback[VAPOR] = feedback[VAPOR] + evap
                feedback[VAPOR] = feedback[VAPOR] + subli
                feedback[HEAT] = feedback[HEAT] - evap * f32(p.evapHeat)
                feedback[HEAT] = feedback[HEAT] - subli * f32(p.evapHeat)
                feedback[HEAT] = feedback[HEAT] - subli * f32(p.meltingHeat)
                new_x = new_x + base[0] / wf * f32(2.0)
                new_y = new_y + base[1] / hf * f32(2.0)
                new_y = new_y - f32(p.fallSpeed) * new_d * np.sqrt(total / area)
                xs = new_x + ONE
                new_x = (xs - f32(2.0) * np.floor(xs / f32(2.0))) - ONE
                feedback[MASS] = total
            surface = f32(12.0) * f32(12.0)
            feedback[MASS] = feedback[MASS] / surface
            feedback[HEAT] = feedback[HEAT] / surface
            feedback[VAPOR] = feedback[VAPOR] / surface
            deposition[0] = deposition[0] / f32(12.0)
            deposition[1] = deposition[1] / f32(12.0)
            size = 12
            gl_x, gl_y = new_x, new_y

        out[n] = (new_x, new_y, new_w, new_i, gmax(new_d, Z))
        splat(gl_x, gl_y, size, np.array(feedback, f32), np.array(deposition, f32))
    return out, fb_tex, dep_tex, seen


@pytest.mark.parametrize("seed,iter_num,cold_cloud", [(3, 205, 30.0), (8, 600, 30.0), (12, 41, 30.0), (20, 77, 7.0)])
def test_precipitation_pass_matches_python_transliteration(seed, iter_num, cold_cloud):
    w, h = 64, 48
    g, base, water, wall, _ = stress_state(w, h, seed=seed)
    g["enablePrecipitation"] = True
    rng = np.random.default_rng(seed)
    # dense, cold cloud aloft: spawning and lightning become likely enough to be exercised
    air = wall[..., 1] != 0
    water[h // 2:, :, 1] += np.where(air[h // 2:], f32(cold_cloud), f32(0.0))
    water[h // 2:, :, 0] += np.where(air[h // 2:], f32(cold_cloud), f32(0.0))
    water[:h // 5, :, 1] += np.where(air[:h // 5], f32(40.0), f32(0.0))   # warm dense cloud near the ground: rain spawns
    water[:h // 5, :, 0] += np.where(air[:h // 5], f32(40.0), f32(0.0))
    n = 400
    drops = np.zeros((n, 5), f32)
    drops[:, 0] = rng.uniform(-1, 1, n)
    drops[:, 1] = rng.uniform(-0.95, 0.95, n)
    drops[:, 2] = rng.uniform(0.0, 0.5, n)
    drops[:, 3] = rng.uniform(0.0, 0.5, n)
    drops[:, 4] = rng.choice([0.2, 0.6, 1.0], n)
    drops[:120, 2] = -2.0 - rng.uniform(-1, 1, 120)           # inactive, position kept as seed
    drops[120:160, 2:4] = rng.uniform(0.0, 0.015, (40, 2))    # residual droplets: evaporate
    drops[160:200, 1] = rng.uniform(-1.0, -0.93, 40)          # in / just above the ground
    drops[200:210, 1] = -1.001                                 # below the map
    drops[210:260, 2] = 0.0                                    # pure ice
    drops[260:300, 3] = 0.0                                    # pure rain
    ora = make_oracle(g, base, water, wall, drops)
    p = P.derive_params(g)
    p.spawnChanceMult = 0.02
    ora.set_params(p)
    # one whole iteration without the particle pass: advection writes the wall markers (TOTAL = 1001 / 1002)
    # the particles test against, and the lighting block toggles `even`, which selects the droplet
    # buffers (app.js:5912-5927)
    for k in range(7):
        ora.run_pass(k)
    ora.iter = iter_num
    ora.inactive_droplets = 3.0
    src = ora.droplets()
    assert np.array_equal(src, drops)
    lightning = ora.lightning
    b1, w1 = ora.field(O.FIELD_BASE, 1), ora.field(O.FIELD_WATER, 1)
    want, want_fb, want_dep, seen = precipitation(b1, w1, src, lightning, p, iter_num, 3.0)
    ora.run_pass(7)
    got = ora.droplets()
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want))).all(axis=1)
    assert not bad.any(), f"{bad.sum()} droplets differ, first {np.argwhere(bad)[0]}: {got[bad][0]} vs {want[bad][0]} from {src[bad][0]}"
    got_fb, got_dep = ora.field(O.FIELD_FEEDBACK), ora.field(O.FIELD_DEPOSITION)
    assert np.array_equal(got_dep, want_dep), f"deposition differs in {(got_dep != want_dep).sum()} values"
    assert np.array_equal(got_fb, want_fb), f"feedback differs in {(got_fb != want_fb).sum()} values"
    # lightningLocationShader.frag:24-38: pixel (1, 0) of the feedback texture is latched unless it holds no strike
    # of this iteration (or two of them, which doubles START_ITERNUM); app.js:5957-5967: the count of inactive
    # droplets in pixel (0, 0) becomes the uniform every 600 iterations
    new = want_fb[0, 1]
    it = f32(iter_num)
    discard = new[2] < gmax(it - ONE, ONE) or new[2] > it
    want_lightning = np.asarray(lightning, f32) if discard else new
    assert np.array_equal(ora.lightning, want_lightning), f"lightning latch {ora.lightning} vs {want_lightning} ({seen['lightning']} spawned)"
    assert ora.inactive_droplets == (want_fb[0, 0, 0] if iter_num % 600 == 0 else 3.0)
    if cold_cloud < 10.0:  # the case chosen to spawn exactly one bolt: it must be latched
        assert seen["lightning"] == 1 and not discard and ora.lightning[2] == want_fb[0, 1, 2] and abs(float(ora.lightning[2]) - (iter_num - 0.15)) < 0.01
    missing = [k for k, v in seen.items() if v == 0 and k not in ("lightning",)]
    assert not missing, f"branches not exercised: {missing} ({seen})"
    print(seen)
