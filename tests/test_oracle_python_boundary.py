"""Independent restatement of the boundary pass (boundaryShader.frag:72-531) as a scalar,
cell-by-cell Python transliteration of the GLSL (numpy float32 scalars: one rounding per
operation; ivec4 wall with the RGBA8I store saturating to [-128, 127]; switch statements with
their fall-through; sin / cos of the uniform sun angle evaluated in double and rounded, pow(x, 17)
as the multiply chain of DESIGN.md 2).  The C++ oracle's boundary pass must reproduce it bit for
bit on a state that visits every wall type, snow, desert, smoke, clouds, precipitation feedback
and deposition, at iteration numbers that trigger the slow processes (every 20 and every 100
iterations).  Like the numpy cross-checks this pins the oracle against transcription slips, not
against the reference's WebGL output."""
import math

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

from util import make_oracle, stress_state

P = wsb200.params
f32 = np.float32
INERT, LAND, WATER, FIRE, URBAN, RUNWAY, INDUSTRIAL = range(7)
VX, VY, PRESSURE, TEMPERATURE = 0, 1, 2, 3
TOTAL, CLOUD, PRECIPITATION, SMOKE = 0, 1, 2, 3
SOIL_MOISTURE, SNOW = 2, 3
TYPE, DISTANCE, VERT_DISTANCE, VEGETATION = 0, 1, 2, 3
SUNLIGHT, NET_HEATING = 0, 1
MASS, HEAT, VAPOR = 0, 1, 2
Z, ONE = f32(0.0), f32(1.0)


def gmax(a, b):
    return b if a < b else a


def gmin(a, b):
    return b if b < a else a


def clamp(x, lo, hi):
    return gmin(gmax(x, lo), hi)


def map_range(v, min1, max1, min2, max2):
    return min2 + (v - min1) * (max2 - min2) / (max1 - min1)


def map_range_c(v, min1, max1, min2, max2):
    return clamp(map_range(v, min1, max1, min2, max2), gmin(min2, max2), gmax(min2, max2))


def max_water(T):
    x = T / f32(250.0)
    x2 = x * x
    x4 = x2 * x2
    x8 = x4 * x4
    x16 = x8 * x8
    return x16 * x


def c_to_k(c):
    return c + f32(273.15)


def boundary(base1, water1, wall1, vort, light0, fb, dep, p, fi, initial_t, iter_num):
    h, w = base1.shape[:2]
    texel_y = f32(1.0 / h)
    sin_s, cos_s = f32(math.sin(float(fi.sunAngle))), f32(math.cos(float(fi.sunAngle)))
    land_evap, water_evap = f32(p.landEvaporation), f32(p.waterEvaporation)
    out_b, out_w, out_wl = np.empty_like(base1), np.empty_like(water1), np.empty_like(wall1)
    it_f = f32(iter_num)
    it_i = int(it_f)

    def calc_evaporation(T, W, V, M):
        return gmax((max_water(T) - W) * land_evap * (V / f32(127.0) + f32(0.1)) * gmin(M + ONE, f32(50.0)) * f32(0.05), Z)

    def calc_fire_intensity(veg, moist, precip):
        return gmax(f32(veg) * f32(0.00025) - moist * f32(0.00020) - precip * f32(0.02), Z)

    for y in range(h):
        yp, ym = (y + 1) % h, (y - 1) % h
        ylp = min(y + 1, h - 1)  # light texture: wrap T = CLAMP_TO_EDGE
        tex_y = (f32(y) + f32(0.5)) * texel_y
        tex_yp = tex_y + texel_y
        for x in range(w):
            xp, xm = (x + 1) % w, (x - 1) % w
            base = [f32(v) for v in base1[y, x]]
            water = [f32(v) for v in water1[y, x]]
            feedback = [f32(v) for v in fb[y, x]]
            real_temp = base[TEMPERATURE] - tex_y * f32(p.dryLapse)
            wall = [int(v) for v in wall1[y, x]]
            w_xm, w_ym = [int(v) for v in wall1[y, xm]], [int(v) for v in wall1[ym, x]]
            w_xp, w_yp = [int(v) for v in wall1[y, xp]], [int(v) for v in wall1[yp, x]]
            light = [f32(v) for v in light0[y, x]]
            next_to_wall = False
            wall[VERT_DISTANCE] = w_ym[VERT_DISTANCE] + 1

            if wall[DISTANCE] != 0:  # fluid
                wall[TYPE] = w_ym[TYPE]
                if wall[TYPE] != WATER:
                    base[TEMPERATURE] = base[TEMPERATURE] + light[NET_HEATING]
                base[TEMPERATURE] = base[TEMPERATURE] + feedback[HEAT]
                coalescence = gmax(-feedback[VAPOR], Z)
                water[CLOUD] = water[CLOUD] - coalescence
                water[TOTAL] = water[TOTAL] - coalescence
                water[TOTAL] = water[TOTAL] + gmax(feedback[VAPOR], Z)
                water[PRECIPITATION] = gmax(water[PRECIPITATION] * f32(0.997) - f32(0.00001) + feedback[MASS] * f32(0.005), Z)
                water[SMOKE] = water[SMOKE] / (ONE + gmax(-feedback[VAPOR] * f32(0.1), Z) + feedback[MASS] * f32(0.000))
                water[SMOKE] = water[SMOKE] - feedback[MASS] * f32(0.0001)
                water[SMOKE] = water[SMOKE] - gmax((water[SMOKE] - f32(4.0)) * f32(0.01), Z)
                water[SMOKE] = gmax(water[SMOKE], Z)
                if water[SMOKE] > f32(4.0):
                    water[SMOKE] = water[SMOKE] - water[PRECIPITATION] * f32(0.02)

                grav = f32(0.0001)
                t_yp = f32(base1[yp, x, TEMPERATURE])
                force = ((base[TEMPERATURE] + t_yp) * f32(0.5) - (f32(initial_t[y]) + f32(initial_t[y + 1])) * f32(0.5)) * grav
                force = force - water[CLOUD] * grav * f32(p.waterWeight)
                force = force - feedback[MASS] * grav * f32(p.waterWeight)
                base[VY] = base[VY] + force

                snow_cover, soil_moisture = Z, Z
                if w_ym[DISTANCE] == 0:
                    next_to_wall = True
                    wall[DISTANCE] = 1
                    snow_cover, soil_moisture = f32(water1[ym, x, SNOW]), f32(water1[ym, x, SOIL_MOISTURE])
                    wall[VERT_DISTANCE] = 1
                if w_xm[DISTANCE] == 0:
                    next_to_wall = True
                    wall[DISTANCE] = 1
                    if w_xm[TYPE] == WATER:
                        wall[TYPE] = LAND
                        wall[DISTANCE] = 0
                    if w_xp[DISTANCE] == 0:
                        wall[DISTANCE] = 0
                elif w_xp[DISTANCE] == 0:
                    next_to_wall = True
                    wall[DISTANCE] = 1
                    if w_xp[TYPE] == WATER:
                        wall[TYPE] = LAND
                        wall[DISTANCE] = 0
                if w_yp[DISTANCE] == 0:
                    next_to_wall = True
                    wall[DISTANCE] = 1
                    if tex_y < f32(0.99):
                        wall[DISTANCE] = 0

                vf, vf_xm, vf_ym = vort[y, x], vort[y, xm], vort[ym, x]
                velocity_factor = np.sqrt(base[VX] * base[VX] + base[VY] * base[VY]) * f32(0.1)
                k = f32(p.vorticity) + velocity_factor
                base[VX] = base[VX] + (f32(vf[0]) + f32(vf_ym[0])) * k
                base[VY] = base[VY] + (f32(vf[1]) + f32(vf_xm[1])) * k

                if next_to_wall:
                    if wall[TYPE] != WATER:
                        power = Z
                        if w_ym[DISTANCE] == 0:
                            power = power + gmax(light[SUNLIGHT] * cos_s, Z)
                        if w_xm[DISTANCE] == 0:
                            power = power + gmax(light[SUNLIGHT] * sin_s, Z)
                        if w_xp[DISTANCE] == 0:
                            power = power + gmax(light[SUNLIGHT] * (-sin_s), Z)
                        albedo = ONE
                        if wall[TYPE] in (LAND, FIRE):
                            soil = map_range_c(soil_moisture, Z, f32(20.0), f32(0.30), f32(0.15))
                            soil = map_range_c(snow_cover, Z, f32(10.0), soil, f32(0.85))
                            full_veg = map_range(snow_cover, Z, f32(10.0), f32(0.10), f32(0.30))
                            albedo = map_range(f32(w_ym[VEGETATION]), Z, f32(127.0), soil, full_veg)
                        elif wall[TYPE] == URBAN:
                            albedo = f32(0.08)
                        elif wall[TYPE] == INDUSTRIAL:
                            albedo = f32(0.08)
                        elif wall[TYPE] == RUNWAY:
                            albedo = f32(0.04)
                        power = power * (ONE - albedo)
                        power = power * f32(0.000002)
                        base[TEMPERATURE] = base[TEMPERATURE] + power
                else:
                    nearest = 255
                    for nb in (w_ym, w_yp, w_xm, w_xp):
                        if nb[DISTANCE] < nearest:
                            nearest = nb[DISTANCE]
                    wall[DISTANCE] = nearest + 1

                if wall[VERT_DISTANCE] <= 5:
                    if wall[VERT_DISTANCE] == 1:
                        drag = f32(0.0015)
                        if wall[TYPE] == URBAN:
                            drag = f32(0.040)
                        elif wall[TYPE] in (LAND, FIRE):
                            drag = map_range_c(f32(wall[VEGETATION]), f32(50.0), f32(127.0), f32(0.0015), f32(0.020))
                        base[VX] = base[VX] - abs(base[VX]) * base[VX] * drag * f32(50.0)
                    rate = f32(0.015)
                    if w_yp[VERT_DISTANCE] <= 5:
                        base[VX] = base[VX] - (base[VX] - f32(base1[yp, x, VX])) * rate
                    if w_ym[VERT_DISTANCE] > 0:
                        base[VX] = base[VX] - (base[VX] - f32(base1[ym, x, VX])) * rate

                if wall[VERT_DISTANCE] <= 8:
                    wall[VEGETATION] = w_ym[VEGETATION]
                    in_surface = [f32(v) for v in water1[ym, x]]
                    t = wall[TYPE]
                    if t == FIRE:
                        if wall[VERT_DISTANCE] == 1:
                            fire = calc_fire_intensity(wall[VEGETATION], in_surface[SOIL_MOISTURE], water[PRECIPITATION])
                            fire = gmax(fire, Z)
                            base[TEMPERATURE] = base[TEMPERATURE] + fire
                            water[SMOKE] = water[SMOKE] + fire * f32(2.0)
                            water[TOTAL] = water[TOTAL] + fire * f32(0.50)
                    if t in (FIRE, INDUSTRIAL):
                        if wall[TYPE] == INDUSTRIAL:
                            tex_frag_x = int(f32(x) + f32(0.5)) % 80
                            if wall[VERT_DISTANCE] == 5 and tex_frag_x in (18, 22):
                                water[TOTAL] = water[TOTAL] + f32(0.25)
                                base[VX] = base[VX] * f32(0.5)
                                base[VY] = base[VY] * f32(0.5)
                                base[VY] = base[VY] + f32(0.05)
                            elif wall[VERT_DISTANCE] == 6 and tex_frag_x == 29:
                                water[SMOKE] = water[SMOKE] + f32(0.01)
                                base[TEMPERATURE] = base[TEMPERATURE] + f32(0.02)
                                base[VX] = base[VX] * f32(0.5)
                                base[VY] = base[VY] * f32(0.5)
                    if t in (FIRE, INDUSTRIAL, URBAN):
                        water[SMOKE] = water[SMOKE] + f32(0.000002)
                    if t in (FIRE, INDUSTRIAL, URBAN, LAND):
                        if wall[VERT_DISTANCE] <= 1:
                            evap = calc_evaporation(real_temp, water[TOTAL], f32(wall[VEGETATION]), in_surface[SOIL_MOISTURE]) / f32(1.0)
                            water[TOTAL] = water[TOTAL] + evap
                            base[TEMPERATURE] = base[TEMPERATURE] - evap * f32(p.evapHeat) * f32(0.5)
                            if wall[VEGETATION] < 10 and water[SOIL_MOISTURE] < f32(5.0):
                                water[SMOKE] = gmin(water[SMOKE] + (gmax(abs(base[VX]) - f32(0.12), Z) * f32(0.15)), f32(2.4))
                    elif t == WATER:
                        if wall[VERT_DISTANCE] <= 1:
                            local_t = f32(base1[ym, x, TEMPERATURE])
                            base[TEMPERATURE] = base[TEMPERATURE] + (local_t - real_temp - ONE) / f32(1.0) * f32(0.0002)
                            water[TOTAL] = water[TOTAL] + gmax((max_water(local_t) - water[TOTAL]) * water_evap / f32(1.0), Z)
            else:  # wall
                wall[VERT_DISTANCE] = w_yp[VERT_DISTANCE] - 1
                if wall[VERT_DISTANCE] < 0:
                    water[2], water[3] = f32(water1[yp, x, 2]), f32(water1[yp, x, 3])
                    wall[VEGETATION] = w_yp[VEGETATION]
                    if w_yp[DISTANCE] == 0:
                        if w_yp[TYPE] != WATER:
                            wall[TYPE] = w_yp[TYPE]
                        elif wall[TYPE] == WATER:
                            base[TEMPERATURE] = f32(base1[yp, x, TEMPERATURE])
                elif wall[VERT_DISTANCE] == 0:
                    water_yp = [f32(v) for v in water1[yp, x]]
                    deposition = [f32(v) for v in dep[y, x]]
                    light_above = [f32(v) for v in light0[ylp, x]]
                    t = wall[TYPE]
                    if t == INDUSTRIAL:
                        wall[VEGETATION] = min(wall[VEGETATION], 15)
                    if t in (INDUSTRIAL, URBAN):
                        wall[VEGETATION] = min(wall[VEGETATION], 75)
                    if t in (INDUSTRIAL, URBAN, FIRE):
                        if wall[TYPE] == FIRE:
                            fire = calc_fire_intensity(wall[VEGETATION], water[SOIL_MOISTURE], water_yp[PRECIPITATION])
                            if fire < f32(0.002):
                                wall[TYPE] = LAND
                            elif it_i % (int(f32(10.0) / fire) + 1) == 0:
                                wall[VEGETATION] -= 1
                                if wall[VEGETATION] < 10:
                                    wall[TYPE] = LAND
                    if t in (INDUSTRIAL, URBAN, FIRE, LAND):
                        water[SOIL_MOISTURE] = clamp(water[SOIL_MOISTURE] + deposition[0] * f32(0.1), Z, f32(1000.0))
                        water[SNOW] = clamp(water[SNOW] + deposition[1] * f32(0.05), Z, f32(4000.0))
                        base_above = [f32(v) for v in base1[yp, x]]
                        real_above = base_above[TEMPERATURE] - tex_yp * f32(p.dryLapse)
                        evap = calc_evaporation(real_above, water_yp[TOTAL], f32(wall[VEGETATION]), water[SOIL_MOISTURE]) * f32(0.10)
                        water[SOIL_MOISTURE] = water[SOIL_MOISTURE] - evap
                        if it_i % 100 == 0:
                            num, tot_snow, tot_soil = Z, Z, Z
                            for nb, xn in ((w_xm, xm), (w_xp, xp)):
                                if nb[VERT_DISTANCE] == 0 and nb[TYPE] in (LAND, URBAN):
                                    tot_snow = tot_snow + f32(water1[y, xn, SNOW])
                                    tot_soil = tot_soil + f32(water1[y, xn, SOIL_MOISTURE])
                                    num = num + ONE
                            if num > Z:
                                water[SNOW] = water[SNOW] + (tot_snow / num - water[SNOW]) * f32(0.02)
                                water[SOIL_MOISTURE] = water[SOIL_MOISTURE] + (tot_soil / num - water[SOIL_MOISTURE]) * f32(0.02)
                            growth = int(water[SOIL_MOISTURE] * np.sqrt(light_above[SUNLIGHT]) * f32(0.01))
                            interval = (100 // growth) * 100 if growth > 0 else 0  # rate > 100: `% 0`, frozen as "no tick"
                            if interval > 0 and it_i % interval == 0:
                                if int(map_range_c(real_above, c_to_k(Z), c_to_k(f32(25.0)), Z, f32(127.0))) > wall[VEGETATION]:
                                    wall[VEGETATION] += 1
                            sub = it_i // 100
                            if (sub % (int(water[SOIL_MOISTURE] * f32(0.1) + water[SNOW] * f32(0.5)) + 10) == 0 and wall[VEGETATION] >= 20
                                    and (w_xm[TYPE] == FIRE or w_xp[TYPE] == FIRE or water_yp[SMOKE] > f32(4.5))):
                                wall[TYPE] = FIRE
                    elif t == WATER:
                        interval = f32(20.0)
                        if f32(p.dynamicWaterTemperature) >= ONE and (it_f - interval * np.floor(it_f / interval)) < f32(0.5):
                            num, tot = Z, Z
                            if w_xm[TYPE] == WATER:
                                tot = tot + f32(base1[y, xm, TEMPERATURE])
                                num = num + ONE
                            if w_xp[TYPE] == WATER:
                                tot = tot + f32(base1[y, xp, TEMPERATURE])
                                num = num + ONE
                            if num > Z:
                                base[TEMPERATURE] = base[TEMPERATURE] + (tot / num - base[TEMPERATURE]) * f32(0.10)
                            if base[TEMPERATURE] > f32(500.0):
                                base[TEMPERATURE] = c_to_k(f32(25.0))
                            air_t = f32(base1[yp, x, TEMPERATURE]) - tex_yp * f32(p.dryLapse)
                            heating = Z
                            heating = heating + (air_t - base[TEMPERATURE]) * f32(0.0002)
                            heating = heating - gmax((max_water(base[TEMPERATURE]) - water_yp[TOTAL]) * water_evap, Z) * f32(p.evapHeat) * f32(0.5)
                            power = gmax(light_above[SUNLIGHT] * cos_s, Z)
                            power = power * (ONE - f32(0.05))
                            power = power * f32(0.000002)
                            heating = heating + power
                            heating = heating + light_above[NET_HEATING]
                            base[TEMPERATURE] = base[TEMPERATURE] + heating / f32(50.0) * interval
                        base[TEMPERATURE] = clamp(base[TEMPERATURE], c_to_k(Z), c_to_k(f32(40.0)))
                        wall[VEGETATION] = 20
                        water[SOIL_MOISTURE] = f32(100.0)
                        water[SNOW] = Z
            out_b[y, x] = base
            out_w[y, x] = water
            out_wl[y, x] = [max(-128, min(127, v)) for v in wall]
    return out_b, out_w, out_wl


# wet=True: soil moisture 250 .. 900 under a high sun pushes vegetationGrowthRate past 100, where the
# reference's growth interval (100 / rate) * 100 is 0 and its `%` undefined (boundaryShader.frag:460-462;
# two shipped saves hold such cells).  Frozen as "no growth tick" in the kernels, the oracle and here.
@pytest.mark.parametrize("iter_num,wet", [(0, False), (7, False), (40, False), (300, False), (300, True), (1000, True)])
def test_boundary_pass_matches_python_transliteration(iter_num, wet):
    w, h = 96, 40
    g, base, water, wall, _ = stress_state(w, h, seed=17)
    if wet:
        g["sunAngle"] = 85.0
        land = (wall[..., 1] == 0) & np.isin(wall[..., 0], (1, 3, 4, 6))  # LAND, FIRE, URBAN, INDUSTRIAL fall through to :460
        water[..., 2] = np.where(land, f32(250.0) + f32(650.0) * np.random.default_rng(5).random((h, w)).astype(f32), water[..., 2])
    g["enablePrecipitation"] = False
    p = P.derive_params(g)
    fi = P.frame_inputs(g)
    initial_t = P.initial_T_profile(h, g)
    ora = make_oracle(g, base, water, wall, None, fi=fi)
    ora.step(12)  # light, wall / vertical distances (one row per iteration: cooling towers at 5, stacks at 6), vegetation settle
    ora.iter = iter_num
    rng = np.random.default_rng(iter_num + 1)
    fb = ora.field(O.FIELD_FEEDBACK, copy=False)
    dep = ora.field(O.FIELD_DEPOSITION, copy=False)
    hit = rng.random((h, w)) < 0.2
    fb[...] = np.where(hit[..., None], rng.normal(0, 0.05, (h, w, 4)), 0).astype(f32)
    fb[..., 0] = np.abs(fb[..., 0]) * 10
    dep[...] = np.where(hit[..., None], rng.uniform(0, 2, (h, w, 2)), 0).astype(f32)
    if wet:  # sunlight travels one row per iteration and has not reached the surface yet: put it there
        ora.field(O.FIELD_LIGHT, 0, copy=False)[..., 0] = f32(900.0)
    ora.run_pass(0)  # velocity -> frameBuff_1
    ora.run_pass(1)  # curl
    ora.run_pass(2)  # vorticity force
    args = (ora.field(O.FIELD_BASE, 1), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 1), ora.field(O.FIELD_VORT),
            ora.field(O.FIELD_LIGHT, 0), fb.copy(), dep.copy(), p, fi, initial_t, iter_num)
    if wet:  # the branch under test must actually be reached
        surf = (args[2][..., 1] == 0) & (np.roll(args[2][..., 1], -1, axis=0) != 0) & np.isin(args[2][..., 0], (1, 3, 4, 6))
        above = np.roll(args[4][..., 0], -1, axis=0)
        assert ((args[1][..., 2] * np.sqrt(above) * f32(0.01))[surf] > 100).sum() > 10
    want_b, want_w, want_wl = boundary(*args)
    ora.run_pass(3)
    got_b, got_w, got_wl = ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 0), ora.field(O.FIELD_WALL, 0)
    bad = (got_wl != want_wl).any(axis=-1)
    assert not bad.any(), f"wall: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got_wl[bad][0]} vs {want_wl[bad][0]} (input {args[2][bad][0]})"
    for name, got, want in (("base", got_b, want_b), ("water", got_w, want_w)):
        for ch in range(4):
            bad = ~((got[..., ch] == want[..., ch]) | (np.isnan(got[..., ch]) & np.isnan(want[..., ch])))
            assert not bad.any(), (f"{name}[{ch}]: {bad.sum()} cells differ, first at {np.argwhere(bad)[0]}: {got[..., ch][bad][0]!r} vs "
                                   f"{want[..., ch][bad][0]!r}; wall in {args[2][bad][0]}")
