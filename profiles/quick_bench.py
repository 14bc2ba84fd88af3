"""Kernel-only timings at the bench grid (no e2e / CPU legs): python profiles/quick_bench.py [dry|full|both] [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
W, H = 16384, 4096
g = P.resolve_settings(None)
g["enablePrecipitation"] = False
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
if mode in ("dry", "both"):
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*wsb200.synth.dry_state(W, H, seed=1234, g=g))
    sim.set_profiling(True)
    sim.step_dry(600)  # SM clock ramp-up after the host-side state generation
    sim.sync()
    sim.step_dry(K)
    t, n = sim.kernel_time_ms(S.KERNEL_DRY)
    print(f"dry: {t / n:.4f} ms/launch  {36 * W * H / (t / n * 1e-3) / 1e9:.0f} GB/s  frac {36 * W * H / (t / n * 1e-3) / 1e9 / 6554.2:.3f}")
    sim.close()
if mode in ("full", "both"):
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    b, w, wl, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False)
    sim.upload(b, w, wl)
    sim.set_profiling(True)
    sim.step(150)
    sim.sync()
    sim.step(K)
    tp, n = sim.kernel_time_ms(S.KERNEL_PVB)
    ta, _ = sim.kernel_time_ms(S.KERNEL_ADV)
    print(f"full: pvb {tp / n:.4f} ms  adv {ta / n:.4f} ms  step {sim.last_step_ms() / K:.4f} ms  {W * H * K / sim.last_step_ms() / 1e6:.2f} Gcell/s")
    sim.close()
