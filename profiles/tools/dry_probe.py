"""Why does the dry sweep's steady-state time differ from a single profiled launch?  Per-launch
times by launch index (state dependence), back to back vs with idle gaps (memory-system state).
    python profiles/tools/dry_probe.py name=lib.so ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
W, H = 16384, 4096
g = P.resolve_settings(None)
state = wsb200.synth.dry_state(W, H, seed=1234, g=g)


def one(sim):
    sim.step_dry(1)
    sim.sync()
    t, n = sim.kernel_time_ms(S.KERNEL_DRY)
    return t / n


for name, path in libs:
    os.environ["WSB200_LIB"] = os.path.abspath(path)
    S._LIB = None
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*state)
    sim.set_profiling(True)
    ts = [one(sim) for _ in range(40)]
    print(name, "single launches, index 1..40:", " ".join(f"{t:.3f}" for t in ts), flush=True)
    sim.step_dry(20)
    sim.sync()
    t, n = sim.kernel_time_ms(S.KERNEL_DRY)
    print(name, f"20 back to back: {t / n:.4f} ms/launch", flush=True)
    ts = []
    for _ in range(8):
        time.sleep(0.02)
        ts.append(one(sim))
    print(name, "single launches after 20 ms idle:", " ".join(f"{t:.3f}" for t in ts), flush=True)
    b = sim.read_pixels(S.FIELD_BASE)
    a = np.abs(b[..., 0:3])
    tiny = np.float32(1.1754944e-38)
    print(name, "denormal fraction vx,vy,P:", [(float(((a[..., k] > 0) & (a[..., k] < tiny)).mean())) for k in range(3)],
          "zero fraction:", [float((a[..., k] == 0).mean()) for k in range(3)], flush=True)
    sim.close()
