"""Process-to-process and allocation-to-allocation spread of the dry sweep's launch time.
    python profiles/tools/dry_probe2.py <lib.so> [W]     (state cached under /dev/shm)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

os.environ["WSB200_LIB"] = os.path.abspath(sys.argv[1])
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
H = 4096
g = P.resolve_settings(None)
cache = f"/dev/shm/dry_{W}x{H}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    state = (z["b"], z["w"], z["wl"])
else:
    state = wsb200.synth.dry_state(W, H, seed=1234, g=g)
    np.savez(cache, b=state[0], w=state[1], wl=state[2])
out = []
for rep in range(3):
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*state)
    sim.set_profiling(True)
    sim.step_dry(30)
    sim.sync()
    sim.step_dry(20)
    sim.sync()
    t, n = sim.kernel_time_ms(S.KERNEL_DRY)
    out.append(t / n * 16384 / W)
    sim.close()
print(os.path.basename(sys.argv[1]), W, "ms/launch (scaled to 16384 columns) for 3 successive allocations:", " ".join(f"{t:.3f}" for t in out), flush=True)
