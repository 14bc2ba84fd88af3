#!/usr/bin/env python
"""bench.py — cell-updates/s of the simulation loop at 16384 x 4096 fp32 on N B200s.

    python bench.py --gpus N --steps K --warmup W                      (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W     (CPU restatement of the
                                                                        reference shaders, host cores)
    N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE iteration of the reference's simulation loop (app.js:5830-6005) over the whole
grid: pressure, velocity, curl, vorticity, boundary, advection (+ condensation), lighting — the
"full physics" workload of BASELINE.json configs[4] (x-strips at 2/4/8 GPUs; the same grid on one
GPU at N = 1).  The total grid is fixed, so scaling is "strong".  Rank 0 prints ONE JSON line.

Timing: CUDA events on the simulation's own stream inside libwsb200 (wsb_last_step_ms) around
exactly K iterations, bracketed by barrier + synchronize, max over ranks.  Inputs are GiB-sized
planes (>> 126 MB L2), so no L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GRID_W, GRID_H = 16384, 4096
METRIC, UNIT = "cell-updates/s", "cell-updates/s"

# algorithmic HBM bytes per cell and launch (DESIGN.md "Kernels and rooflines")
B_ALG = {
    "k_fused_pvb": 80,   # R base 16 + wall 4 + water 16 + light (sun, net heating) 8; W base 16 + water 16 + wall 4
    "k_fused_adv": 100,  # R base 16 + water 16 + wall 4 + light (sun, IR down, IR up) 12; W base 16 + water 16 + wall 4 + light 16
    "k_fused_dry": 36,   # R base 16 + wall 4; W base 16
}
PREWARM_ITERS = 150      # untimed, before the W warm-up steps
B_ALG_STEP_FULL = 104    # every live field read once + written once per iteration


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while a timed region runs.
    NVML is initialised in the constructor (it can take longer than a 50 ms timed region); the
    thread then samples every 5 ms until result() is called."""

    NAMES = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.reasons, self.sm_max, self.err, self.nv, self.h = [], set(), None, None, None, None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: fall back to one nvidia-smi sample
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for k, bit in self.NAMES.items():
                if mask & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        if self.nv is None:
            return
        try:
            while not self.stop_flag.is_set():
                self._sample()
                time.sleep(0.005)
        except Exception as e:
            self.err = repr(e)

    def result(self):
        if self.nv is not None and self.is_alive():
            try:
                self._sample()  # one more while the GPU is still under load / just finished
            except Exception:
                pass
        self.stop_flag.set()
        self.join(timeout=2)
        if not self.sm:
            try:
                import subprocess

                out = subprocess.check_output(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                               "--format=csv,noheader,nounits"], text=True).strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "note": f"single nvidia-smi sample after the run ({self.err})"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": f"no clock source: {self.err}"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def _ncu_traffic(kernel: str, W: int, H: int, world: int):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json),
    only when the capture was taken at this grid on one GPU; otherwise null."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t["grid"] == [W, H] and world == 1:
            return t[kernel]["read"] + t[kernel]["write"]
    except Exception:
        pass
    return None


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle, timed on the host cores (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def _oracle_sample(width_cols: int, height: int):
    """A periodic slab of the bench state: `width_cols` columns x full height."""
    import wsb200
    from oracle import oracle as O

    P = wsb200.params
    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    base, water, wall, _ = wsb200.synth.full_state(GRID_W, height, seed=7, g=g, with_droplets=False, cols=np.arange(width_cols))
    O.set_threads()  # all host cores (torchrun exports OMP_NUM_THREADS=1)
    ora = O.OracleSim(width_cols, height, 0)
    ora.upload(base, water, wall, None)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(height, g))
    return ora


def cpu_baseline(budget_s: float = 12.0):
    cols, h = 512, GRID_H
    ora = _oracle_sample(cols, h)
    ora.step(1)
    t = time.perf_counter()
    ora.step(2)
    per = (time.perf_counter() - t) / 2
    n = int(max(3, min(400, budget_s / max(per, 1e-6))))
    t = time.perf_counter()
    ora.step(n)
    dt = time.perf_counter() - t
    from oracle import oracle as O

    cores = O.lib().oracle_get_threads()
    return {"value": cols * h * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cols}x{h} periodic slab of the bench state (full physics, no particles), {n} iterations, OpenMP over rows; "
                      "CPU restatement of the reference shaders (oracle/wsb_oracle.cpp) — the reference itself (GLSL under a browser) cannot run here"}


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    cols, h = 512, GRID_H
    ora = _oracle_sample(cols, h)
    ora.step(max(args.warmup, 1))
    t = time.perf_counter()
    ora.step(args.steps)
    dt = time.perf_counter() - t
    value = cols * h * args.steps / dt
    from oracle import oracle as O

    cores = O.lib().oracle_get_threads()
    sample = (f"each step = one iteration on a {cols}x{h} periodic slab of the 16384x4096 bench state; value = slab cells x steps / time; "
              "oracle/wsb_oracle.cpp (C++/OpenMP restatement of the reference shaders; the GLSL/browser reference cannot be executed on this box)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "full physics 16384x4096 fp32, no particles (CPU: bounded slab sample)", "grid": [GRID_W, GRID_H]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def _max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world: int):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def _pinned(shape, dtype):
    import torch

    tdt = {np.float32: torch.float32, np.int8: torch.int8}[dtype]
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()


def run_ours(args):
    import torch

    import wsb200

    S, P = wsb200.sim, wsb200.params
    rank, local_rank, world = _dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    W, H = args.width, args.height
    K, Wm = args.steps, max(args.warmup, 3)
    peak, peak_src = _peaks()

    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    sim = wsb200.multi.create_distributed(W, H, device=local_rank, gui_controls=g)
    x0, lw, gh = sim.layout()
    cols = sim.padded_columns()
    base, water, wall, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False, cols=cols)
    hb, hw, hl = _pinned(base.shape, np.float32), _pinned(water.shape, np.float32), _pinned(wall.shape, np.int8)
    hb[...], hw[...], hl[...] = base, water, wall
    del base, water, wall
    sim.upload_local(hb, hw, hl)
    sim.set_profiling(True)

    # ---- device-timed leg: inputs resident in HBM -------------------------------------------
    # The GPU idles while the host generates the state; run until the SM clock has ramped up again
    # (untimed), then the W warm-up steps proper.
    if not args.no_prewarm:  # a FIXED count: every rank must run the same number of halo exchanges
        sim.step(PREWARM_ITERS)
        sim.sync()
    sim.step(Wm)
    sim.sync()
    _barrier(world)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sim.launch_count
    sim.step(K)
    sim.sync()
    ms = sim.last_step_ms()
    _barrier(world)
    clocks = sampler.result()
    launches = sim.launch_count - launches0
    kt = {name: sim.kernel_time_ms(k) for name, k in (("k_fused_pvb", S.KERNEL_PVB), ("k_fused_adv", S.KERNEL_ADV), ("halo", S.KERNEL_HALO))}
    ms = _max_over_ranks(ms, world, device)
    value = W * H * K / (ms * 1e-3)
    vmax = sim.max_velocity

    # roofline of the dominant kernel (this rank's strip; cells include the ghost columns it computes)
    local_cells = (lw + 2 * gh) * H
    dom = max(("k_fused_pvb", "k_fused_adv"), key=lambda n: kt[n][0])
    dom_ms = kt[dom][0] / max(kt[dom][1], 1)
    achieved = B_ALG[dom] * local_cells / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(dom, W, H, world), "peak_source": peak_src, "alg_bytes_per_cell": B_ALG[dom], "avg_launch_ms": dom_ms,
                "kernels_ms_per_step": {n: (t / max(c, 1)) for n, (t, c) in kt.items()},
                "step_frac_of_104B_roofline": (B_ALG_STEP_FULL * W * H / world / (ms / K * 1e-3) / 1e9) / peak}

    # ---- end-to-end leg: host buffers in, host buffers out, through the public API ------------
    ob, ow, ol = _pinned((H, lw, 4), np.float32), _pinned((H, lw, 4), np.float32), _pinned((H, lw, 4), np.int8)
    fi = sim.frame_inputs
    _barrier(world)
    t0 = time.perf_counter()
    sim.upload_local(hb, hw, hl)                      # loadData -> textures (pinned host -> HBM)
    for _ in range(K):
        sim.set_frame_inputs(fi)                      # per-frame uniforms
        sim.step(1)
    sim.read_pixels(S.FIELD_BASE, x0, 0, lw, H, out=ob)   # prepareDownload: frameBuff_0 readback
    sim.read_pixels(S.FIELD_WATER, x0, 0, lw, H, out=ow)
    sim.read_pixels(S.FIELD_WALL, x0, 0, lw, H, out=ol)
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, device)
    h2d = (hb.nbytes + hw.nbytes + hl.nbytes) * world / K + 56
    d2h = (ob.nbytes + ow.nbytes + ol.nbytes) * world / K
    e2e = {"value": W * H * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "what": f"upload of the state from pinned host memory + {K} x (set_frame_inputs + step(1)) + readback of base/water/wall to pinned host, "
                   "host wall clock, bytes amortised over the K steps"}
    finite = bool(np.isfinite(ob).all())
    sim.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full physics {W}x{H} fp32 (pressure+velocity+vorticity+boundary+advection+condensation+lighting), no particles",
                       "grid": [W, H], "partition": f"{world} x-strip(s) of {lw} columns, ghost {gh}, one NCCL ring exchange per iteration" if world > 1 else "single GPU",
                       "schedule": "fused: k_fused_pvb + k_fused_adv per iteration (TMA-staged channel planes)", "prewarm": f"{PREWARM_ITERS} untimed iterations before the W warm-up steps (SM clock ramp-up after host-side state generation)", "l2": "no flush: every plane is >= 256 MiB, far larger than the 126 MB L2",
                       "max_abs_velocity_cells_per_iter": vmax, "state_finite": finite},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    if world == 1 and rank == 0:
        line["dry_sweep"] = dry_sweep_leg(W, H, K, Wm, peak, local_rank, prewarm=not args.no_prewarm)
        if args.particles:
            line["with_particles"] = particles_leg(W, H, K, Wm, local_rank)
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def dry_sweep_leg(W, H, K, Wm, peak, device_index, prewarm=True):
    """The fused pressure+velocity+advection sweep (k_fused_dry) at the same grid: the kernel the
    >= 70 % HBM-roofline target of BASELINE.json is stated on."""
    import wsb200

    S, P = wsb200.sim, wsb200.params
    g = P.resolve_settings(None)
    sim = wsb200.Simulation(W, H, 0, device=device_index, gui_controls=g)
    base, water, wall = wsb200.synth.dry_state(W, H, seed=1234, g=g)
    sim.upload(base, water, wall)
    del base, water, wall
    sim.set_profiling(True)
    if prewarm:
        sim.step_dry(4 * PREWARM_ITERS)  # clock ramp-up after the host-side state generation
        sim.sync()
    sim.step_dry(Wm)
    sim.sync()
    sampler = ClockSampler(device_index)
    sampler.start()
    sim.step_dry(K)
    sim.sync()
    clocks = sampler.result()
    ms = sim.last_step_ms()
    t, c = sim.kernel_time_ms(S.KERNEL_DRY)
    per = t / max(c, 1)
    achieved = B_ALG["k_fused_dry"] * W * H / (per * 1e-3) / 1e9
    out = {"value": W * H * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
           "roofline": {"bound": "hbm", "kernel": "k_fused_dry", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": _ncu_traffic("k_fused_dry", W, H, 1), "alg_bytes_per_cell": B_ALG["k_fused_dry"], "avg_launch_ms": per},
           "max_abs_velocity_cells_per_iter": sim.max_velocity, "clocks": clocks}
    sim.close()
    return out


def particles_leg(W, H, K, Wm, device_index):
    """BASELINE config 4: full physics + 1 M precipitation particles on one GPU."""
    import wsb200

    S, P = wsb200.sim, wsb200.params
    g = P.resolve_settings(None)
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    nd = 1_000_000
    sim = wsb200.Simulation(W, H, nd, device=device_index, gui_controls=g)
    base, water, wall, drops = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=True, n_droplets=nd)
    wsb200.synth.add_clouds(base, water, wall, n_blobs=96, seed=5)  # something to rain from
    sim.upload(base, water, wall, drops)
    del base, water, wall
    sim.set_profiling(True)
    sim.step(max(Wm, 30))  # spin-up: the first iterations spawn the bulk of the droplets
    sim.sync()
    sim.step(K)
    sim.sync()
    ms = sim.last_step_ms()
    kt = {n: sim.kernel_time_ms(k) for n, k in (("k_fused_pvb", S.KERNEL_PVB), ("k_fused_adv", S.KERNEL_ADV), ("k_precipitation", S.KERNEL_PRECIP))}
    d = sim.read_droplets()
    out = {"value": W * H * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "n_droplets": nd, "active_droplets": int((d[:, 2] >= 0).sum()),
           "kernels_ms_per_step": {n: t / max(c, 1) for n, (t, c) in kt.items()}}
    sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=GRID_W)
    ap.add_argument("--height", type=int, default=GRID_H)
    ap.add_argument("--particles", action="store_true", help="also run BASELINE config 4 (1 M droplets) at N=1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-prewarm", action="store_true", help="skip the clock ramp-up iterations (profiler runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
