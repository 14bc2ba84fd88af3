"""BASELINE config 4 timing only: python profiles/quick_particles.py [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W, H, nd = 16384, 4096, 1_000_000
g = P.resolve_settings(None)
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
sim = wsb200.Simulation(W, H, nd, gui_controls=g)
base, water, wall, drops = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=True, n_droplets=nd)
wsb200.synth.add_clouds(base, water, wall, n_blobs=96, seed=5)
sim.upload(base, water, wall, drops)
sim.set_profiling(True)
sim.step(60)
sim.step(K)
kt = {n: sim.kernel_time_ms(k) for n, k in (("pvb", S.KERNEL_PVB), ("adv", S.KERNEL_ADV), ("precip", S.KERNEL_PRECIP))}
d = sim.read_droplets()
print(f"step {sim.last_step_ms() / K:.4f} ms  " + "  ".join(f"{n} {t / max(c, 1):.4f}" for n, (t, c) in kt.items()) + f"  active {(d[:, 2] >= 0).sum()}")
