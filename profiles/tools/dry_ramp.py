"""Dry-sweep time per launch over consecutive batches (clock ramp-up / steady state), with the SM
clock sampled by nvidia-smi next to each batch:  python profiles/tools/dry_ramp.py name=lib.so ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
W, H = 16384, 4096
g = P.resolve_settings(None)
g["enablePrecipitation"] = False
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
state = wsb200.synth.dry_state(W, H, seed=1234, g=g)


def clocks():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                              capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return repr(e)


for name, path in libs:
    os.environ["WSB200_LIB"] = os.path.abspath(path)
    S._LIB = None
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*state)
    sim.set_profiling(True)
    for batch in range(12):
        sim.step_dry(100)
        sim.sync()
        t, n = sim.kernel_time_ms(S.KERNEL_DRY)
        print(f"{name} batch {batch:2d}: {t / n:.4f} ms/launch  wall {sim.last_step_ms() / 100:.4f} ms/iter  maxv {sim.max_velocity:.3f} | {clocks()}", flush=True)
    sim.close()
