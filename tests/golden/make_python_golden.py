"""Generates tests/golden/python_100x100.npz: the state of the reference's own save
`100x100_test.weathersandbox` (saves/100 X 100 Test.weathersandbox) after 1, 10 and 101 iterations of
the full loop INCLUDING the precipitation particles, computed by the independent Python
restatements of the shaders (tests/test_oracle_numpy_*.py, tests/test_oracle_python_*.py: numpy
fp32 / scalar np.float32 transliterations written from the GLSL) chained in the order of
app.js:5830-6005 — NOT by the C++ oracle and NOT by the CUDA path.  tests/test_python_golden.py
holds the oracle and the fused kernels (on the host emulator) to these vectors bit for bit.

Still not the reference's WebGL output (parity unpinned): it is a second reading of
the same sources, with the frozen choices of DESIGN.md 2.

    python tests/golden/make_python_golden.py        (about a minute)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import wsb200  # noqa: E402
from test_oracle_numpy_advection import _advection  # noqa: E402
from test_oracle_numpy_crosscheck import _pressure, _tex, _velocity  # noqa: E402
from test_oracle_numpy_lighting import _lighting  # noqa: E402
from test_oracle_python_boundary import boundary  # noqa: E402
from test_oracle_python_precipitation import gmax, precipitation  # noqa: E402

P = wsb200.params
f32 = np.float32


def curl_and_vorticity(base):
    vx, vy = base[..., 0], base[..., 1]
    curl = ((_tex(vx, 0, 1) - vx) - _tex(vy, 1, 0)) + vy
    ac = np.abs(curl)
    fx = _tex(ac, 0, -1) - _tex(ac, 0, 1)
    fy = _tex(ac, 1, 0) - _tex(ac, -1, 0)
    mag = np.sqrt(fx * fx + fy * fy) + f32(0.0001)
    return np.stack([(fx / mag) * curl, (fy / mag) * curl], axis=-1)


class PythonSim:
    """app.js:5830-6005 with the bindings of every pass (which ping-pong copy plays which role)."""

    def __init__(self, sf, g):
        self.g, self.p, self.fi = g, P.derive_params(g), P.frame_inputs(g)
        h, w = sf.height, sf.width
        self.initial_t = P.initial_T_profile(h, g)
        self.zeros = np.zeros(h + 1, f32)
        self.base = [sf.base.copy(), sf.base.copy()]
        self.water = [sf.water.copy(), sf.water.copy()]
        self.wall = [sf.wall.copy(), sf.wall.copy()]
        self.light = [np.zeros((h, w, 4), f32), np.zeros((h, w, 4), f32)]
        self.fb, self.dep = np.zeros((h, w, 4), f32), np.zeros((h, w, 2), f32)
        self.drops = sf.droplets.copy()
        self.lightning = np.zeros(4, f32)
        self.inactive = 0.0
        self.even, self.iter = True, 0

    def iteration(self):
        p, fi = self.p, self.fi
        self.base[1] = _velocity(self.base[0], self.wall[0], p.dragMultiplier, p.wind)
        self.wall[1] = self.wall[0].copy()
        vort = curl_and_vorticity(self.base[1])
        self.base[0], self.water[0], self.wall[0] = boundary(self.base[1], self.water[1], self.wall[1], vort, self.light[0], self.fb, self.dep,
                                                             p, fi, self.initial_t, self.iter)
        self.base[1], self.water[1], self.wall[1] = _advection(self.base[0], self.water[0], self.wall[0], p, self.zeros, self.zeros, self.zeros)
        self.base[0] = _pressure(self.base[1], self.wall[1])
        self.wall[0] = self.wall[1].copy()
        src, dst = (0, 1) if self.even else (1, 0)
        self.light[dst] = _lighting(self.base[1], self.water[1], self.wall[1], self.light[src], p, fi)
        self.even = not self.even
        if p.enablePrecipitation:
            self.drops, self.fb, self.dep, _ = precipitation(self.base[1], self.water[1], self.drops, self.lightning, p, self.iter, self.inactive)
            if self.iter % 600 == 0:
                self.inactive = float(self.fb[0, 0, 0])
            new, it = self.fb[0, 1], f32(self.iter)
            if not (new[2] < gmax(it - f32(1.0), f32(1.0)) or new[2] > it):
                self.lightning = new.copy()
        self.iter += 1

    def light_latest(self):
        return self.light[0 if self.even else 1]


def run(iters=(1, 10, 101)):
    sf = wsb200.savefile.load(os.path.join(HERE, "100x100_test.weathersandbox"))
    g = P.resolve_settings(sf.settings_json)
    sim = PythonSim(sf, g)
    out, done = {}, 0
    for n in iters:
        for _ in range(n - done):
            sim.iteration()
        done = n
        out[f"base_{n}"], out[f"water_{n}"], out[f"wall_{n}"] = sim.base[0].copy(), sim.water[1].copy(), sim.wall[0].copy()
        out[f"light_{n}"], out[f"drops_{n}"] = sim.light_latest().copy(), sim.drops.copy()
        out[f"feedback_{n}"], out[f"deposition_{n}"] = sim.fb.copy(), sim.dep.copy()
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "python_100x100.npz")
    np.savez_compressed(path, **run())
    print("written", path)
