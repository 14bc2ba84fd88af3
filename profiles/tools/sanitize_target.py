"""Small driver for compute-sanitizer: the two smoke configurations, a few iterations each.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python profiles/tools/sanitize_target.py [iters]
100 x 100 reference save with particles (register-staged edge tiles, sprite atomics, latches) and a
320 x 128 new-simulation state (TMA-staged interior tiles, in-place shared-memory sweeps), plus the
dry sweep and the readback kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wsb200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S, P = wsb200.sim, wsb200.params
sf = wsb200.savefile.load(os.path.join(ROOT, "tests", "golden", "100x100_test.weathersandbox"))
sim = wsb200.Simulation.from_save(sf)
sim.step(n)
sim.read_pixels(S.FIELD_BASE)
sim.read_pixels(S.FIELD_WALL)
sim.read_droplets()
sim.close()

g = P.resolve_settings(None)
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
w, h = 320, 128
base, water, wall, drops = wsb200.synth.setup_state(w, h, seed=0.61, height_mult=0.8, g=g, with_droplets=True)
base[..., 0] += np.where(wall[..., 1] != 0, np.float32(0.05), np.float32(0))
sim = wsb200.Simulation(w, h, drops.shape[0], gui_controls=g)
sim.upload(base, water, wall, drops)
sim.step(n)
sim.read_pixels(S.FIELD_BASE)
sim.read_pixels(S.FIELD_LIGHT, view=S.VIEW_LATEST)
sim.read_points(S.FIELD_BASE, np.array([[5, 5], [300, 100]], dtype=np.int32))
sim.close()

g["enablePrecipitation"] = False
sim = wsb200.Simulation(w, h, 0, gui_controls=g)
sim.upload(*wsb200.synth.dry_state(w, h, seed=3, g=g))
sim.step_dry(n)
sim.read_pixels(S.FIELD_BASE)
sim.close()
print("sanitize target done", n)
