"""BASELINE config 4 timing only: python profiles/quick_particles.py [K] [n_droplets]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W, H, nd = 16384, 4096, (int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000)
g = P.resolve_settings(None)
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
sim = wsb200.Simulation(W, H, nd, gui_controls=g)
base, water, wall, drops = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=True, n_droplets=nd)
wsb200.synth.add_clouds(base, water, wall, n_blobs=96, seed=5)
sim.upload(base, water, wall, drops)
sim.set_profiling(True)
sim.step(180)
sim.step(K)
kt = {n: sim.kernel_time_ms(k) for n, k in (("pvb", S.KERNEL_PVB), ("adv", S.KERNEL_ADV), ("particle pass", S.KERNEL_PRECIP), ("of which boxsum+clear", S.KERNEL_SPRITES))}
d = sim.read_droplets()
fb = sim.read_pixels(S.FIELD_FEEDBACK)
tiles = (fb[..., :3] != 0).any(axis=-1).reshape(H // 16, 16, W // 64, 64).any(axis=(1, 3))
print(f"feedback texels hit {(fb[..., :3] != 0).any(axis=-1).mean():.4f} of the grid; 64x16 tiles with a hit {tiles.mean():.4f}")
print(f"step {sim.last_step_ms() / K:.4f} ms  " + "  ".join(f"{n} {t / max(c, 1):.4f}" for n, (t, c) in kt.items()) + f"  active {(d[:, 2] >= 0).sum()}")
