import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, wsb200
from util import stress_state, make_cuda
S = wsb200.sim
g, base, water, wall, _ = stress_state(192, 96, seed=7)
g["enablePrecipitation"] = False
sim = make_cuda(g, base, water, wall, None, S.SCHEDULE_FUSED)
mode = sys.argv[1]
try:
    if mode == "dry": sim.step_dry(1)
    else: sim.step(1)
    sim.sync()
    print("ok", mode)
except Exception as e:
    print("FAIL", mode, e)
