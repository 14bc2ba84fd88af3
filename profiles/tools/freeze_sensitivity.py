#!/usr/bin/env python
"""How much room does "implementation-defined" leave?  The reference's own shaders (oracle/_ref, compiled from the
checkout) are built three ways and run from the reference's 100 x 100 save (BASELINE config 1):

  canonical   the spec freeze of DESIGN.md 2 (closed-form pow, sin / cos through double, no contraction) — the build the
              oracle and the kernels reproduce bit for bit
  libm        pow -> powf, sin / cos -> sinf / cosf: what an implementation with correctly-rounded-ish library
              transcendentals would compute
  fma         canonical built with -ffp-contract=fast: what a compiler that contracts a*b+c (every desktop GLSL
              compiler may) would compute

and the relative distance (|a - b| / (|b| + 1e-3), max over the field) of each to the canonical build is printed
after 1, 10, 100, 1000 iterations.  CPU only; needs the reference checkout.

    python profiles/tools/freeze_sensitivity.py > profiles/r3_freeze_sensitivity.log
"""
import ctypes
import importlib.util
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import wsb200  # noqa: E402
from oracle import ref_shaders as R  # noqa: E402

P = wsb200.params


def variant_sim(flags, sf, g, tmp, name):
    """A RefShaderSim backed by an experiment build of the library."""
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "oracle", "ref_shim", "build_ref.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    path = b.build("/root/reference", extra_flags=flags, out=os.path.join(tmp, f"libref_{name}.so"))
    saved = R._lib
    R._lib = None
    orig_build = R.build
    R.build = lambda force=False: path
    try:
        R.lib()
        sim = R.RefShaderSim(sf.width, sf.height, sf.droplets.shape[0])
    finally:
        R.build = orig_build
        R._lib = saved
    sim.upload(sf.base, sf.water, sf.wall, sf.droplets)
    sim.set_params(P.derive_params(g))
    sim.set_frame_inputs(P.frame_inputs(g))
    sim.set_profiles(P.initial_T_profile(sf.height, g))
    return sim


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3)))


def main():
    sf = wsb200.savefile.load(os.path.join(ROOT, "tests", "golden", "100x100_test.weathersandbox"))
    g = P.resolve_settings(sf.settings_json)
    with tempfile.TemporaryDirectory() as tmp:
        sims = {"canonical": variant_sim((), sf, g, tmp, "canonical"),
                "libm": variant_sim(("-DWSB_REF_LIBM",), sf, g, tmp, "libm"),
                "fma": variant_sim(("-ffp-contract=fast", "-march=x86-64-v3"), sf, g, tmp, "fma")}
        print("reference shaders on saves/100 X 100 Test: distance of two other legal implementations to the canonical build")
        print(f"{'iterations':>10} {'variant':>9} {'base':>10} {'water':>10} {'light':>10} {'wall bytes':>10}")
        done = 0
        for n in (1, 10, 100, 1000):
            for s in sims.values():
                s.step(n - done)
            done = n
            c = sims["canonical"]
            for name in ("libm", "fma"):
                s = sims[name]
                print(f"{n:>10} {name:>9} {rel(s.field(0, 0), c.field(0, 0)):>10.3g} {rel(s.field(1, 1), c.field(1, 1)):>10.3g} "
                      f"{rel(s.light_latest(), c.light_latest()):>10.3g} {int((s.field(2, 0) != c.field(2, 0)).sum()):>10}")


if __name__ == "__main__":
    main()
