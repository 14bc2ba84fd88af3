"""Host-side parameter derivation (app.js:347-407, 3394-3443, 5439-5474, 6510-6572)."""
import json
import math

import numpy as np

import wsb200

P = wsb200.params


def test_old_save_defaults(save100):
    g = P.resolve_settings(save100.settings_json)
    # keys the shipped saves predate: numeric -> default, boolean -> false
    assert g["condensationRate"] == 0.005
    assert g["globalEffectsStartAlt"] == 0 and g["globalEffectsEndAlt"] == 10000
    assert g["soundingForcing"] == 0.0
    assert g["dynamicWaterTemperature"] is False
    assert g["sound"] is False
    # keys the save has win over the defaults
    assert g["evapHeat"] == 1.9 and g["meltingHeat"] == 0.6 and g["waterWeight"] == 0.5
    assert g["vorticity"] == 0.007 and g["dragMultiplier"] == 0.01 and g["wind"] == -0.0001


def test_no_save_is_defaults():
    g = P.resolve_settings(None, sim_height=8000)
    assert g["simHeight"] == 8000 and g["globalEffectsEndAlt"] == 8000
    assert g["evapHeat"] == 2.90 and g["dynamicWaterTemperature"] is True


def test_minus_one_is_replaced():
    g = P.resolve_settings(json.dumps({"evapHeat": -1, "wind": -0.5}))
    assert g["evapHeat"] == 2.90 and g["wind"] == -0.5


def test_derive_params_uniform_values(save100):
    g = P.resolve_settings(save100.settings_json)
    p = P.derive_params(g)
    assert p.dryLapse == np.float32(120.0)  # 12000 m * 10 K/km
    assert p.globalEffectsStartAlt == 0.0
    assert abs(p.globalEffectsEndAlt - 10000 / 12000) < 1e-7
    assert abs(p.waterTemperature - (g["waterTemperature"] + 273.15)) < 1e-4
    assert p.dynamicWaterTemperature == 0.0
    assert p.enablePrecipitation == 1
    assert p.spawnChanceMult == np.float32(g["spawnChance"])
    assert ctypes_sizeof(p) == 32 * 4


def ctypes_sizeof(x):
    import ctypes

    return ctypes.sizeof(x)


def test_initial_T_profile():
    g = P.resolve_settings(None)
    t = P.initial_T_profile(300, g)
    assert t.shape == (301,) and t.dtype == np.float32
    assert abs(t[0] - 288.15) < 1e-4  # 15 C at the ground
    # app.js:5469-5474 at y = 150, H = 300
    alt = 150 / 301 * 12000
    want = max(15 + alt * (-85) / 12000, -60) + 273.15 + (150 / 300) * 120.0
    assert abs(t[150] - want) < 1e-3
    # the real-temperature floor of -60 C is reached near the top
    assert abs((t[300] - 120.0) - 213.15) < 1e-3
    # generalisation beyond the reference's 504 entries
    t2 = P.initial_T_profile(4096, g)
    assert t2.shape == (4097,) and np.all(np.diff(t2) > 0)


def test_sun_uniforms():
    zen, inten = P.sun_uniforms(90.0, 1.0)  # overhead
    assert zen == 0.0 and abs(inten - 1300.0) < 1e-6
    zen, inten = P.sun_uniforms(60.0, 1.0)
    assert abs(zen - (-30 * P.DEG_TO_RAD)) < 1e-12
    assert abs(inten - math.sin(120 * P.DEG_TO_RAD) ** 0.1 * 1300.0) < 1e-9
    _, night = P.sun_uniforms(200.0, 1.0)
    assert night == 0.0


def test_frame_inputs_idle(save100):
    g = P.resolve_settings(save100.settings_json)
    fi = P.frame_inputs(g)
    assert fi.userInputType == -1 and list(fi.userInputValues) == [0, 0, 0, 0] and list(fi.airplaneValues) == [0, 0, 0, 0]
    assert abs(fi.sunAngle - (g["sunAngle"] - 90) * P.DEG_TO_RAD) < 1e-6


def test_sun_clock_advances():
    g = P.resolve_settings(None)
    g["timeOfDay"], g["month"], g["latitude"] = 12.0, 6.65, 45.0
    clk = P.SunClock(g)
    a0 = g["sunAngle"]
    clk.advance(1.0)
    assert abs(g["timeOfDay"] - 13.0) < 0.02
    assert g["sunAngle"] != a0
    z, i = clk.uniforms()
    assert i > 1000.0
