"""A/B kernel timings of several builds of libwsb200 in ONE process (states are generated once):

    python profiles/tools/ab_bench.py [--small] [--k K] name=path/to/lib.so [name=...]

For every library: the dry sweep and the full-physics step at the bench grid (16384 x 4096, or
4096 x 1024 with --small), per-kernel times from the library's own CUDA-event timers, and a
checksum of the state after the run — builds that claim bit-exact parity must print the same one."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import wsb200  # noqa: E402

S, P = wsb200.sim, wsb200.params
args = sys.argv[1:]
small = "--small" in args
K = int(args[args.index("--k") + 1]) if "--k" in args else 20
libs = [a.split("=", 1) for a in args if "=" in a]
W, H = (4096, 1024) if small else (16384, 4096)
PEAK = 6458.7
g = P.resolve_settings(None)
g["enablePrecipitation"] = False
g["dayNightCycle"] = False
g["sunAngle"] = 60.0
dry_state = wsb200.synth.dry_state(W, H, seed=1234, g=g)
fb, fw, fwl, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False)
print(f"grid {W}x{H}, K={K}", flush=True)


def digest(sim, field, view):
    return hashlib.sha1(sim.read_pixels(field, view=view).tobytes()).hexdigest()[:12]


for name, path in libs:
    os.environ["WSB200_LIB"] = os.path.abspath(path)
    S._LIB = None
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(*dry_state)
    sim.set_profiling(True)
    sim.step_dry(300 if not small else 50)  # SM clock ramp-up
    sim.sync()
    sim.step_dry(K)
    t, n = sim.kernel_time_ms(S.KERNEL_DRY)
    ms = t / n
    print(f"{name:10s} dry : {ms:.4f} ms/launch  {36 * W * H / (ms * 1e-3) / 1e9:7.0f} GB/s  frac {36 * W * H / (ms * 1e-3) / 1e9 / PEAK:.3f}"
          f"  sha {digest(sim, S.FIELD_BASE, 0)}", flush=True)
    sim.close()
    sim = wsb200.Simulation(W, H, 0, gui_controls=g)
    sim.upload(fb, fw, fwl)
    sim.set_profiling(True)
    sim.step(100 if not small else 20)
    sim.sync()
    sim.step(K)
    tp, n = sim.kernel_time_ms(S.KERNEL_PVB)
    ta, _ = sim.kernel_time_ms(S.KERNEL_ADV)
    print(f"{name:10s} full: pvb {tp / n:.4f} ms  adv {ta / n:.4f} ms  step {sim.last_step_ms() / K:.4f} ms  "
          f"{W * H * K / sim.last_step_ms() / 1e6:.2f} Gcell/s  sha {digest(sim, S.FIELD_BASE, 0)} {digest(sim, S.FIELD_WATER, 1)}", flush=True)
    sim.close()
