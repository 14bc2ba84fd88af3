"""Small driver for ncu: a few iterations of each fused kernel at the bench grid.
    ncu ... python profiles/prof_target.py [full|dry] [W H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wsb200  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "full"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
H = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
P = wsb200.params
g = P.resolve_settings(None)
g["enablePrecipitation"] = False
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
if mode == "dry":
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*wsb200.synth.dry_state(W, H, seed=1234, g=g))
    sim.step_dry(n)
else:
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    b, w, wl, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False)
    sim.upload(b, w, wl)
    sim.step(n)
sim.sync()
print("done", mode, W, H, sim.launch_count)
