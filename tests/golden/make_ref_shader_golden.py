"""Generates tests/golden/ref_shader_golden.npz: states produced by the REFERENCE'S OWN SHADERS, executed on the CPU
through oracle/_ref/libref_shaders.so (oracle/ref_shim/: the GLSL files of the reference checkout, translated
mechanically and compiled against glsl_shim.h; the draw loop of app.js:5830-6005 restated in ref_driver.cpp).

These are the vectors the GPU box — which has no reference checkout — checks the CUDA path (and the oracle) against
(tests/test_ref_shader_golden.py).  Needs /root/reference:

    python tests/golden/make_ref_shader_golden.py            # rewrite the .npz
    python tests/golden/make_ref_shader_golden.py --check    # compare with the committed file, exit 1 on a difference

Cases (inputs come from files committed under tests/golden/ or from tests/util.stress_state, so nothing but the
outputs needs storing):

  save100        the reference's `saves/100 X 100 Test` (BASELINE config 1), parameters resolved like its loader,
                 precipitation as the save says (on): 1, 10, 100 iterations; base / water / wall / droplets also
                 after 1000
  save100_dry    the same with precipitation off: 1, 10, 100, 1000
  stress         128 x 64 synthetic stress state (every wall type, fire, snow, desert, smoke, clouds; power-of-two
                 grid, so the light is exact too), precipitation off: 1, 10, 100
  stress_rain    the same with its droplets, precipitation on: 1, 100
  hotlake        the 128 x 128 window (columns 192.., rows from the ground) of hotlake_valley_crop512 — a state the
                 reference's WebGL path wrote — precipitation off: 100
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import wsb200  # noqa: E402
from oracle import ref_shaders as R  # noqa: E402

P = wsb200.params
OUT = os.path.join(HERE, "ref_shader_golden.npz")
HOTLAKE_WINDOW = (192, 128, 128)  # first column, width, height


def case_inputs(name):
    """(g, base, water, wall, droplets) of a case; used by the generator and by the tests."""
    from util import stress_state

    if name.startswith("save100"):
        sf = wsb200.savefile.load(os.path.join(HERE, "100x100_test.weathersandbox"))
        g = P.resolve_settings(sf.settings_json)
        if name == "save100_dry":
            g["enablePrecipitation"] = False
        return g, sf.base, sf.water, sf.wall, sf.droplets
    if name.startswith("stress"):
        g, base, water, wall, drops = stress_state(128, 64, seed=29)
        g["enablePrecipitation"] = name == "stress_rain"
        return g, base, water, wall, drops
    if name == "hotlake":
        sf = wsb200.savefile.load(os.path.join(HERE, "hotlake_valley_crop512.weathersandbox"))
        x0, w, h = HOTLAKE_WINDOW
        g = P.resolve_settings(sf.settings_json)
        g["enablePrecipitation"] = False
        cut = lambda a: np.ascontiguousarray(a[:h, x0:x0 + w])  # noqa: E731
        return g, cut(sf.base), cut(sf.water), cut(sf.wall), None
    raise KeyError(name)


CASES = {"save100": (1, 10, 100, 1000), "save100_dry": (1, 10, 100, 1000), "stress": (1, 10, 100), "stress_rain": (1, 100), "hotlake": (100,)}


def run():
    out = {}
    for name, snaps in CASES.items():
        g, base, water, wall, drops = case_inputs(name)
        h, w = base.shape[:2]
        use_drops = drops is not None and g["enablePrecipitation"]
        ref = R.RefShaderSim(w, h, drops.shape[0] if use_drops else 0)
        ref.upload(base, water, wall, drops if use_drops else None)
        ref.set_params(P.derive_params(g))
        ref.set_frame_inputs(P.frame_inputs(g))
        ref.set_profiles(P.initial_T_profile(h, g))
        done = 0
        for n in snaps:
            ref.step(n - done)
            done = n
            out[f"{name}/base/{n}"] = ref.field(R.FIELD_BASE, 0)     # frameBuff_0: after the pressure pass
            out[f"{name}/water/{n}"] = ref.field(R.FIELD_WATER, 1)   # frameBuff_1: after advection
            out[f"{name}/wall/{n}"] = ref.field(R.FIELD_WALL, 0)
            if n < 1000:
                out[f"{name}/light/{n}"] = ref.light_latest()
            if use_drops:
                out[f"{name}/drops/{n}"] = ref.droplets()
                out[f"{name}/lightning/{n}"] = ref.lightning
        ref.close()
    return out


if __name__ == "__main__":
    got = run()
    if "--check" in sys.argv:
        want = np.load(OUT)
        bad = [k for k in got if k not in want or not np.array_equal(got[k], want[k], equal_nan=True)]
        print("ref_shader_golden.npz:", "up to date" if not bad else f"{len(bad)} arrays differ: {bad[:5]}")
        sys.exit(1 if bad else 0)
    np.savez_compressed(OUT, **got)
    print("written", OUT, f"({os.path.getsize(OUT) / 1e6:.1f} MB, {len(got)} arrays)")
