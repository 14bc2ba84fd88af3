"""How far does a contraction-allowed build drift from the frozen (no-FMA) arithmetic?  
GLSL ES leaves the precision of mix() to the implementation; desktop GPUs contract a*(1-t)+b*t into FMAs.  This runs
the -DWSB_EXP_FMAMIX build of libwsb200 (NOT the shipped arithmetic) against the oracle on the reference's 100 x 100
save and on a stress state and prints the largest relative error after 1 / 10 / 100 / 1000 iterations.

    make -C 2d-weather-sandbox_b200/csrc variant OUT=$PWD/gpurun_in/libwsb200_fma.so EXTRA=-DWSB_EXP_FMAMIX
    python profiles/tools/fma_tolerance.py gpurun_in/libwsb200_fma.so"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["WSB200_LIB"] = os.path.abspath(sys.argv[1])
import numpy as np  # noqa: E402

import wsb200  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import make_cuda, make_oracle, rel_err, stress_state  # noqa: E402

S, P = wsb200.sim, wsb200.params
print(wsb200.load_library().wsb_build_info().decode())
sf = wsb200.savefile.load(os.path.join(ROOT, "tests", "golden", "100x100_test.weathersandbox"))
cases = {"100x100 save": (P.resolve_settings(sf.settings_json), sf.base, sf.water, sf.wall, sf.droplets)}
g, b, w, wl, d = stress_state(192, 96, seed=3)
cases["stress 192x96"] = (g, b, w, wl, d)
for name, (g, b, w, wl, d) in cases.items():
    for particles in (False, True):
        g = dict(g)
        g["enablePrecipitation"] = particles
        sim = make_cuda(g, b, w, wl, d if particles else None, S.SCHEDULE_FUSED)
        ora = make_oracle(g, b, w, wl, d if particles else None)
        done = 0
        for n in (1, 10, 100, 1000):
            sim.step(n - done)
            ora.step(n - done)
            done = n
            wall_same = np.array_equal(sim.read_pixels(S.FIELD_WALL), ora.field(O.FIELD_WALL, 0))
            e = {"base": rel_err(sim.read_pixels(S.FIELD_BASE), ora.field(O.FIELD_BASE, 0)),
                 "water": rel_err(sim.read_pixels(S.FIELD_WATER, view=1), ora.field(O.FIELD_WATER, 1)),
                 "light": rel_err(sim.read_pixels(S.FIELD_LIGHT, view=S.VIEW_LATEST), ora.light_latest())}
            print(f"{name:14s} particles={particles!s:5s} n={n:4d}  wall identical: {wall_same}  max rel err " +
                  "  ".join(f"{k} {v:.3g}" for k, v in e.items()), flush=True)
        sim.close()
