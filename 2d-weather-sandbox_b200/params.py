"""Host-side parameter derivation for the simulation loop.

Restates, in Python, what the reference computes in JavaScript before/around the hot loop:

* `guiControls_default`                      app.js:347-407
* settings loading semantics                  dat.gui.min.js:135-151 + app.js:3394-3398
  (numeric keys missing from an old save load as -1 and are replaced by the default,
  missing booleans load as false)
* the uniform values of setGuiUniforms        app.js:3401-3443
* dryLapse and the initial_T profile          app.js:5439, 5467-5474, 708
* the sun model updateSunlight                app.js:6495-6572, 3886-3911

Everything is computed in double precision like JS and rounded to float32 when it is stored in
the C-ABI structs (a WebGL `uniform1f` rounds the same way).
"""
from __future__ import annotations

import ctypes
import datetime as _dt
import json
import math

import numpy as np

# --------------------------------------------------------------------------------------------
# C-ABI structs (include/wsb200.h)
# --------------------------------------------------------------------------------------------
_PARAM_FLOATS = [
    "dragMultiplier", "wind", "vorticity", "landEvaporation", "waterEvaporation",
    "dynamicWaterTemperature", "evapHeat", "waterWeight", "meltingHeat", "condensationRate",
    "globalDrying", "globalHeating", "soundingForcing", "globalEffectsStartAlt",
    "globalEffectsEndAlt", "waterTemperature", "greenhouseGases", "waterGreenHouseEffect",
    "IR_rate", "dryLapse", "aboveZeroThreshold", "subZeroThreshold", "spawnChanceMult",
    "snowDensity", "fallSpeed", "growthRate0C", "growthRate_30C", "freezingRate", "meltingRate",
    "evapRate",
]


class WsbParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in _PARAM_FLOATS] + [
        ("enablePrecipitation", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class WsbFrameInputs(ctypes.Structure):
    _fields_ = [
        ("sunAngle", ctypes.c_float),
        ("sunIntensity", ctypes.c_float),
        ("userInputValues", ctypes.c_float * 4),
        ("userInputMove", ctypes.c_float * 2),
        ("userInputType", ctypes.c_int32),
        ("wrapHorizontally", ctypes.c_int32),
        ("airplaneValues", ctypes.c_float * 4),
    ]


# --------------------------------------------------------------------------------------------
# guiControls
# --------------------------------------------------------------------------------------------
GUI_DEFAULTS = {  # app.js:347-407
    "vorticity": 0.005, "dragMultiplier": 0.001, "wind": 0.0, "globalEffectsStartAlt": 0,
    "globalEffectsEndAlt": 10000, "globalDrying": 0.0, "globalHeating": 0.0,
    "soundingForcing": 0.0, "sunIntensity": 1.0, "waterTemperature": 25.0,
    "dynamicWaterTemperature": True, "landEvaporation": 0.00005, "waterEvaporation": 0.0001,
    "evapHeat": 2.90, "meltingHeat": 0.43, "condensationRate": 0.0050, "waterWeight": 0.25,
    "inactiveDroplets": 0, "aboveZeroThreshold": 1.0, "subZeroThreshold": 0.005,
    "spawnChance": 0.00005, "snowDensity": 0.2, "fallSpeed": 0.0003, "growthRate0C": 0.0001,
    "growthRate_30C": 0.001, "freezingRate": 0.01, "meltingRate": 0.01, "evapRate": 0.0008,
    "displayMode": "DISP_REAL", "wrapHorizontally": True, "SmoothCam": True, "camSpeed": 0.01,
    "exposure": 1.0, "timeOfDay": 9.9, "latitude": 45.0, "month": 6.65, "sunAngle": 9.9,
    "dayNightCycle": True, "greenhouseGases": 0.001, "waterGreenHouseEffect": 0.0015,
    "IR_rate": 1.0, "tool": "TOOL_NONE", "brushSize": 20, "wholeWidth": False,
    "intensity": 0.01, "showGraph": False, "realDewPoint": False, "enablePrecipitation": True,
    "showDrops": False, "paused": False, "IterPerFrame": 10, "auto_IterPerFrame": True,
    "sound": True, "dryLapseRate": 10.0, "simHeight": 12000, "twelveHourClock": False,
    "lengthUnit": "LENGTH_UNIT_METRIC", "tempUnit": "TEMP_UNIT_C", "windUnit": "SPEED_UNIT_KMH",
}

TIME_PER_ITERATION = 0.00008  # hours, app.js:449


def resolve_settings(settings_json: str | None, sim_height: float | None = None) -> dict:
    """Effective guiControls after loading.

    No save file (app.js:3378-3391): the defaults, with simHeight / globalEffectsEndAlt taken from
    the new-simulation dialog.  Save file (app.js:3392-3399): keys present in the JSON win; a
    numeric key the save does not have is created as -1 by the patched dat.GUI and then replaced
    by its default; a missing boolean is created as false; a missing selector takes its first
    option (not relevant to the simulation loop).
    """
    if settings_json is None:
        g = dict(GUI_DEFAULTS)
        if sim_height is not None:
            g["simHeight"] = sim_height
            g["globalEffectsEndAlt"] = sim_height
        return g
    loaded = json.loads(settings_json)
    g = {}
    for key, default in GUI_DEFAULTS.items():
        if key in loaded:
            g[key] = loaded[key]
        elif isinstance(default, bool):
            g[key] = False
        else:
            g[key] = default  # -1 -> default (numbers); selectors are UI only
    for key, value in loaded.items():  # stale keys stay in the object and are re-saved
        g.setdefault(key, value)
    # a genuine -1 stored in a save is also replaced (app.js:3395: `value === -1`)
    for key, value in list(g.items()):
        if value == -1 and not isinstance(value, bool) and key in GUI_DEFAULTS:
            g[key] = GUI_DEFAULTS[key]
    return g


def c_to_k(c: float) -> float:
    return c + 273.15


def map_range(value, low1, high1, low2, high2):  # app.js:516
    return low2 + ((high2 - low2) * (value - low1)) / (high1 - low1)


def dry_lapse(g: dict) -> float:
    """app.js:5439"""
    return (g["simHeight"] * g["dryLapseRate"]) / 1000.0


def initial_T_profile(height: int, g: dict) -> np.ndarray:
    """app.js:5467-5474 generalised from 504 entries to height+1."""
    lapse = dry_lapse(g)
    out = np.zeros(height + 1, np.float32)
    for y in range(height + 1):
        altitude = y / (height + 1) * g["simHeight"]
        real_temp = max(map_range(altitude, 0, 12000, 15.0, -70.0), -60)
        out[y] = c_to_k(real_temp) + (y / height) * lapse  # realToPotentialT, app.js:708
    return out


def derive_params(g: dict) -> WsbParams:
    """setGuiUniforms (app.js:3401-3443) + static uniforms (app.js:5478-5640)."""
    p = WsbParams()
    p.dragMultiplier = g["dragMultiplier"]
    p.wind = g["wind"]
    p.vorticity = g["vorticity"]
    p.landEvaporation = g["landEvaporation"]
    p.waterEvaporation = g["waterEvaporation"]
    p.dynamicWaterTemperature = 1.0 if g["dynamicWaterTemperature"] else 0.0
    p.evapHeat = g["evapHeat"]
    p.waterWeight = g["waterWeight"]
    p.meltingHeat = g["meltingHeat"]
    p.condensationRate = g["condensationRate"]
    p.globalDrying = g["globalDrying"]
    p.globalHeating = g["globalHeating"]
    p.soundingForcing = g["soundingForcing"]
    p.globalEffectsStartAlt = g["globalEffectsStartAlt"] / g["simHeight"]
    p.globalEffectsEndAlt = g["globalEffectsEndAlt"] / g["simHeight"]
    p.waterTemperature = c_to_k(g["waterTemperature"])
    p.greenhouseGases = g["greenhouseGases"]
    p.waterGreenHouseEffect = g["waterGreenHouseEffect"]
    p.IR_rate = g["IR_rate"]
    p.dryLapse = dry_lapse(g)
    p.aboveZeroThreshold = g["aboveZeroThreshold"]
    p.subZeroThreshold = g["subZeroThreshold"]
    p.spawnChanceMult = g["spawnChance"]
    p.snowDensity = g["snowDensity"]
    p.fallSpeed = g["fallSpeed"]
    p.growthRate0C = g["growthRate0C"]
    p.growthRate_30C = g["growthRate_30C"]
    p.freezingRate = g["freezingRate"]
    p.meltingRate = g["meltingRate"]
    p.evapRate = g["evapRate"]
    p.enablePrecipitation = 1 if g["enablePrecipitation"] else 0
    return p


# --------------------------------------------------------------------------------------------
# sun model
# --------------------------------------------------------------------------------------------
DEG_TO_RAD = 0.0174533  # app.js:340
RAD_TO_DEG = 57.2957795  # app.js:341


def sun_uniforms(sun_angle_deg: float, sun_intensity_gui: float) -> tuple[float, float]:
    """The uniform part of updateSunlight (app.js:6538-6550): returns
    (solarZenithAngle [rad], sunIntensity [W/m2])."""
    zenith = (sun_angle_deg - 90) * DEG_TO_RAD
    intensity = sun_intensity_gui * math.pow(max(math.sin((180.0 - sun_angle_deg) * DEG_TO_RAD), 0.0), 0.1) * 1300.0
    return zenith, intensity


class SunClock:
    """Date-driven part of updateSunlight (app.js:6510-6536) plus its initialisation in
    startSimulation / onUpdateTimeOfDaySlider / onUpdateMonthSlider (app.js:3902-3910, 6495-6508).
    JS `Date` local-time arithmetic is restated with a naive datetime (no DST, no time zone)."""

    def __init__(self, g: dict):
        self.g = g
        month = g["month"]
        # new Date(2000, floor(month) - 1, (month % 1) * 30.417): day-of-month truncates, day 0 is
        # the last day of the previous month
        self.t = _dt.datetime(2000, 1, 1) + _dt.timedelta(days=0)
        self.t = self._set_month(_dt.datetime(2000, 1, 1), math.floor(month) - 1, int((month % 1) * 30.417))
        if g["dayNightCycle"]:
            tod = g["timeOfDay"]
            minutes = int((tod % 1) * 60)
            self.t = self.t.replace(hour=0, minute=0) + _dt.timedelta(hours=int(tod), minutes=minutes)
            self._recompute()
            m = g["month"] - 0.96
            self.t = self._set_month(self.t, int(m), int((m % 1) * 30))
            self._recompute()

    @staticmethod
    def _set_month(t: _dt.datetime, month_index: int, day: int) -> _dt.datetime:
        year = t.year + month_index // 12
        first = t.replace(year=year, month=month_index % 12 + 1, day=1)
        return first + _dt.timedelta(days=day - 1)

    def _recompute(self):
        g = self.g
        tod_rad = (g["timeOfDay"] / 24.0) * 2.0 * math.pi - math.pi / 2.0
        tilt_deg = math.sin(g["month"] * 0.5236 - 1.92) * 23.5
        t = tilt_deg * DEG_TO_RAD
        lat = g["latitude"] * DEG_TO_RAD
        ang = math.asin(math.sin(t) * math.sin(lat) + math.cos(t) * math.cos(lat) * math.sin(tod_rad)) * RAD_TO_DEG
        if g["latitude"] - tilt_deg < 0.0:
            ang = 180.0 - ang
        g["sunAngle"] = ang

    def advance(self, hours: float):
        """updateSunlight(deltaT_hours): once per FRAME (app.js:5815-5821)."""
        ms = int(hours * 3600 * 1000)  # Date(ms) truncates to whole milliseconds
        self.t = self.t + _dt.timedelta(milliseconds=ms)
        g = self.g
        g["timeOfDay"] = self.t.hour + self.t.minute / 60.0 + self.t.second / 3600.0
        g["month"] = (self.t.month - 1) + 1 + self.t.day / 30.5 + self.t.hour / 720.0
        self._recompute()

    def uniforms(self) -> tuple[float, float]:
        return sun_uniforms(self.g["sunAngle"], self.g["sunIntensity"])


def frame_inputs(g: dict, sun: tuple[float, float] | None = None) -> WsbFrameInputs:
    """Idle-frame inputs: no brush (userInputType = -1, app.js:5750,5808; the other brush uniforms
    keep their GL default 0 until the first mouse press, including `wrapHorizontally`,
    app.js:5804-5806), airplane block inert (airplaneValues = 0)."""
    zen, inten = sun if sun is not None else sun_uniforms(g["sunAngle"], g["sunIntensity"])
    fi = WsbFrameInputs()
    fi.sunAngle = zen
    fi.sunIntensity = inten
    fi.userInputType = -1
    fi.wrapHorizontally = 0
    return fi
