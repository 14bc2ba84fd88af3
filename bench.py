#!/usr/bin/env python
"""bench.py — cell-updates/s of the simulation loop at 16384 x 4096 fp32 on N B200s.

    python bench.py --gpus N --steps K --warmup W                      (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference's shaders compiled for the
                                                                        host, oracle/_ref — else the oracle port — host cores)
    N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE iteration of the reference's simulation loop (app.js:5830-6005) over the whole
grid: pressure, velocity, curl, vorticity, boundary, advection (+ condensation), lighting — the
"full physics" workload of BASELINE.json configs[4] (x-strips at 2/4/8 GPUs; the same grid on one
GPU at N = 1).  The total grid is fixed, so scaling is "strong".  Rank 0 prints ONE JSON line.

Timing: CUDA events on the simulation's own stream inside libwsb200 (wsb_last_step_ms) around
exactly K iterations, bracketed by barrier + synchronize, max over ranks (`value`); `sustained` is
the median over SUSTAINED_BATCHES further batches of K iterations each, timed the same way.  Inputs
are GiB-sized planes (>> 126 MB L2), so no L2 flush is needed between iterations.

At N = 1 the line also carries every other BASELINE.json config as its own object with `roofline`
and `clocks`: `dry_sweep` (k_fused_dry at the headline grid: the >= 70 % target), `config2_dry_4096x1024`,
`config3_full_8192x2048`, `with_particles` (config 4: + 1 M droplets) and `with_particles_ref_count`
(the W*H/25 = 2.68 M droplets the reference itself would allocate, app.js:452,1282).
At N > 1 rank 0 re-runs the end-to-end leg on ONE GPU and compares every strip bit for bit
(`config.bit_identical_to_1gpu`); a flow beyond the strips' ghost budget fails the run (rc != 0).
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
SUSTAINED_BATCHES = 10   # further timed batches of K steps (median reported beside the single batch)
E2E_LONG_K = 200         # second end-to-end figure: PCIe traffic amortised over 200 steps (N = 1)
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


TRAFFIC_SOURCE = "profiles/ncu_traffic.json (static: ncu --set full capture of this kernel at this grid, not measured in this run)"


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


def _p2p_attrs(dev: int, peers):
    """cudaDeviceGetP2PAttribute of this rank's device towards its ring neighbours: performance rank, access, native
    atomics (1 over NVLink, 0 over PCIe) — recorded because the peer transport's 32-byte stores crawl over PCIe P2P."""
    import ctypes

    out = {}
    try:
        rt = None
        for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
            try:
                rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if rt is None:
            import glob

            import torch

            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
            rt = ctypes.CDLL(cands[0])
        for p in sorted(set(peers)):
            if p == dev:
                continue
            vals = {}
            for attr, key in ((1, "performance_rank"), (2, "access"), (3, "native_atomics")):
                v = ctypes.c_int(-1)
                rc = rt.cudaDeviceGetP2PAttribute(ctypes.byref(v), attr, dev, p)
                vals[key] = v.value if rc == 0 else f"error {rc}"
            out[f"{dev}->{p}"] = vals
    except Exception as e:  # diagnostics only
        out["error"] = repr(e)
    return out


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# CPU side, timed on the host cores (cpu_baseline leg and --impl reference): the reference's own shaders compiled
# for the host (oracle/_ref/libref_shaders.so, kind "reference") where that library exists — it is built where the
# reference checkout is and travels to the GPU box with the snapshot — else the oracle port (kind "port")
# ------------------------------------------------------------------------------------------------
CPU_SLAB_COLS = 512


def _slab_state(width_cols: int, height: int):
    """A periodic slab of the bench state: `width_cols` columns x full height."""
    import wsb200

    P = wsb200.params
    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    base, water, wall, _ = wsb200.synth.full_state(GRID_W, height, seed=7, g=g, with_droplets=False, cols=np.arange(width_cols))
    return g, base, water, wall


def _cpu_sims(width_cols: int, height: int):
    """[(kind, sim, cores, what)]: the compiled reference shaders first when available, then the port."""
    import wsb200
    from oracle import oracle as O

    P = wsb200.params
    g, base, water, wall = _slab_state(width_cols, height)
    cores = O.set_threads()  # all host cores (torchrun exports OMP_NUM_THREADS=1)
    sims = []
    try:
        from oracle import ref_shaders as R

        if R.available():
            R.lib().refsim_set_threads(cores)
            sims.append(("reference", R.RefShaderSim(width_cols, height, 0), cores,
                         "the reference's own GLSL simulation shaders, translated mechanically and compiled for the host "
                         "(oracle/_ref/libref_shaders.so, oracle/ref_shim/), one OpenMP thread per core over rows"))
    except Exception as e:  # the library is optional: fall back to the port, loudly
        print(f"bench: oracle/_ref not usable ({e}); CPU arm = oracle port", file=sys.stderr)
    sims.append(("port", O.OracleSim(width_cols, height, 0), cores,
                 "oracle/wsb_oracle.cpp, the C++/OpenMP restatement of the reference shaders (bit-identical to them, tests/test_ref_shaders.py)"))
    for _, sim, _, _ in sims:
        sim.upload(base, water, wall, None)
        sim.set_params(P.derive_params(g))
        sim.set_frame_inputs(P.frame_inputs(g))
        sim.set_profiles(P.initial_T_profile(height, g))
    return sims


def _time_cpu(sim, budget_s: float, cells: int):
    sim.step(1)
    t = time.perf_counter()
    sim.step(2)
    per = (time.perf_counter() - t) / 2
    n = int(max(3, min(400, budget_s / max(per, 1e-6))))
    t = time.perf_counter()
    sim.step(n)
    return cells * n / (time.perf_counter() - t), n


def cpu_baseline(budget_s: float = 12.0):
    cols, h = CPU_SLAB_COLS, GRID_H
    out = None
    for kind, sim, cores, what in _cpu_sims(cols, h):
        value, n = _time_cpu(sim, budget_s if out is None else budget_s / 3, cols * h)
        entry = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "same_config": False,
                 "workload": f"per-cell rate extrapolated from a {cols}x{h} slab = 1/{GRID_W // cols} of the 16384x4096 grid",
                 "sample": f"{cols}x{h} periodic slab of the bench state (full physics, no particles), {n} iterations; {what}"}
        if out is None:
            out = entry
        else:  # the faster CPU implementation of the same arithmetic, for scale
            out["port"] = {k: entry[k] for k in ("value", "unit", "cores", "sample")}
        sim.close()
    return out


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    cols, h = CPU_SLAB_COLS, GRID_H
    kind, sim, cores, what = _cpu_sims(cols, h)[0]
    sim.step(max(args.warmup, 1))
    t = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t
    value = cols * h * args.steps / dt
    sample = f"each step = one iteration on a {cols}x{h} periodic slab of the 16384x4096 bench state; value = slab cells x steps / time; {what}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full physics 16384x4096 fp32, no particles — CPU arm: per-cell rate extrapolated from a {cols}x{h} slab = 1/{GRID_W // cols} of the grid per step",
                       "grid": [GRID_W, GRID_H], "same_config": False},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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


def _digest(*arrays) -> str:
    import hashlib

    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _timed_batches(sim, step, K, n, world, device):
    """n further batches of K steps, each bracketed like the headline batch; ms per step of each (max over ranks)."""
    out = []
    for _ in range(n):
        _barrier(world)
        step(K)
        sim.sync()
        out.append(_max_over_ranks(sim.last_step_ms(), world, device) / K)
    return out


def _roofline(kernel, per_launch_ms, cells, peak, peak_src, W, H, world):
    achieved = B_ALG[kernel] * cells / (per_launch_ms * 1e-3) / 1e9
    t = _ncu_traffic(kernel, W, H, world)
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": t, "traffic_source": TRAFFIC_SOURCE if t is not None else None, "peak_source": peak_src,
            "alg_bytes_per_cell": B_ALG[kernel], "avg_launch_ms": per_launch_ms}


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
    transport = os.environ.get("WSB_EXCHANGE", "auto")

    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    sim = wsb200.multi.create_distributed(W, H, device=local_rank, gui_controls=g, transport=transport)
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
    calibration = None
    if world > 1 and transport == "auto":  # both transports are set up: time them on the live state, keep the faster (untimed)
        calibration = wsb200.multi.calibrate_exchange(sim)
    if world > 1:
        transport = getattr(sim, "transport", transport)
    sim.step(Wm)
    sim.sync()
    sampler = ClockSampler(local_rank)  # NVML initialisation takes tens of ms, differently on every rank: BEFORE the barrier —
    _barrier(world)                     # a rank that starts late makes its neighbours wait inside their timed region
    sampler.start()
    launches0 = sim.launch_count
    sim.step(K)
    sim.sync()
    ms = sim.last_step_ms()
    _barrier(world)
    clocks = sampler.result()
    launches = sim.launch_count - launches0
    kt = {name: sim.kernel_time_ms(k) for name, k in (("k_fused_pvb", S.KERNEL_PVB), ("k_fused_adv", S.KERNEL_ADV), ("halo", S.KERNEL_HALO),
                                                       ("ghost_wait", S.KERNEL_WAIT), ("edge_tiles", S.KERNEL_EDGE))}
    ms_mine = ms
    ms = _max_over_ranks(ms, world, device)
    value = W * H * K / (ms * 1e-3)
    vmax = sim.max_velocity

    # roofline of the dominant kernel (this rank's strip; cells include the ghost columns it computes).
    # On a strip a kernel class is several launches per iteration (interior + edge tile columns): per-iteration time.
    local_cells = (lw + 2 * gh) * H
    dom = max(("k_fused_pvb", "k_fused_adv"), key=lambda n: kt[n][0])
    dom_ms = kt[dom][0] / K
    roofline = _roofline(dom, dom_ms, local_cells, peak, peak_src, W, H, world)
    roofline["kernels_ms_per_step"] = {n: t / K for n, (t, c) in kt.items()}
    roofline["launches_per_step"] = {n: c / K for n, (t, c) in kt.items()}
    roofline["step_frac_of_104B_roofline"] = (B_ALG_STEP_FULL * W * H / world / (ms / K * 1e-3) / 1e9) / peak
    if world > 1:  # every rank's own view: strips run in loose lock-step, the slowest one sets the pace
        import torch.distributed as dist

        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "ms_per_step": ms_mine / K,
                                          "p2p": _p2p_attrs(local_rank, [(local_rank - 1) % world, (local_rank + 1) % world]), "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"),
                                          **{n: round(t / K, 4) for n, (t, c) in kt.items()}})
        roofline["per_rank"] = per_rank

    # ---- sustained: the median of further batches of K steps ---------------------------------
    sus_sampler = ClockSampler(local_rank)
    _barrier(world)
    sus_sampler.start()
    batches = _timed_batches(sim, sim.step, K, 1 if args.no_prewarm else (3 if args.quick else SUSTAINED_BATCHES), world, device)
    sus_clocks = sus_sampler.result()
    med = float(np.median(batches))
    sustained = {"batches": len(batches), "steps_per_batch": K, "ms_per_step_median": med, "ms_per_step_min": min(batches), "ms_per_step_max": max(batches),
                 "value_median": W * H / (med * 1e-3), "unit": UNIT, "clocks": sus_clocks}

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "ms_per_step": ms / K, "sustained": sustained, "transport": transport,
                              "transport_calibration_ms_per_step": calibration, "per_rank": roofline.get("per_rank"),
                              "kernels_ms_per_step": roofline["kernels_ms_per_step"]}), flush=True)
        sim.close()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end-to-end leg: host buffers in, host buffers out, through the public API ------------
    ob, ow, ol = _pinned((H, lw, 4), np.float32), _pinned((H, lw, 4), np.float32), _pinned((H, lw, 4), np.int8)
    fi = sim.frame_inputs

    def e2e_run(k):
        _barrier(world)
        t0 = time.perf_counter()
        sim.upload_local(hb, hw, hl)                      # loadData -> textures (pinned host -> HBM)
        for _ in range(k):
            sim.set_frame_inputs(fi)                      # per-frame uniforms
            sim.step(1)
        sim.read_pixels(S.FIELD_BASE, x0, 0, lw, H, out=ob)   # prepareDownload: frameBuff_0 readback
        sim.read_pixels(S.FIELD_WATER, x0, 0, lw, H, out=ow)
        sim.read_pixels(S.FIELD_WALL, x0, 0, lw, H, out=ol)
        torch.cuda.synchronize()
        return _max_over_ranks(time.perf_counter() - t0, world, device)

    e2e_s = e2e_run(K)
    h2d_total = (hb.nbytes + hw.nbytes + hl.nbytes) * world
    d2h_total = (ob.nbytes + ow.nbytes + ol.nbytes) * world
    e2e = {"value": W * H * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_total / K + 56, "d2h_bytes_per_step": d2h_total / K, "steps": K,
           "seconds": e2e_s,
           "what": f"upload of the state from pinned host memory + {K} x (set_frame_inputs + step(1)) + readback of base/water/wall to pinned host, "
                   f"host wall clock; the {h2d_total / 1e9:.2f} GB up + {d2h_total / 1e9:.2f} GB down are amortised over these {K} steps, so the figure moves with K"}
    finite = bool(np.isfinite(ob).all())

    # ---- N > 1: the strips against ONE GPU, bit for bit (same upload + K steps as the end-to-end leg) ----
    ident = None
    if world > 1:
        import torch.distributed as dist

        light = sim.read_pixels(S.FIELD_LIGHT, x0, 0, lw, H, view=S.VIEW_LATEST)
        mine = {"x0": x0, "lw": lw, "base": _digest(ob), "water": _digest(ow), "wall": _digest(ol), "light": _digest(light)}
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            fb_, fw_, fl_, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False)
            one = wsb200.Simulation(W, H, 0, device=local_rank, gui_controls=g)
            one.upload(fb_, fw_, fl_)
            del fb_, fw_, fl_
            for _ in range(K):
                one.set_frame_inputs(fi)
                one.step(1)
            full = {"base": one.read_pixels(S.FIELD_BASE), "water": one.read_pixels(S.FIELD_WATER), "wall": one.read_pixels(S.FIELD_WALL),
                    "light": one.read_pixels(S.FIELD_LIGHT, view=S.VIEW_LATEST)}
            one.close()
            bad = [f"rank {r}: {name}" for r, p_ in enumerate(parts) for name in ("base", "water", "wall", "light")
                   if _digest(full[name][:, p_["x0"]:p_["x0"] + p_["lw"]]) != p_[name]]
            ident = {"ok": not bad, "mismatches": bad, "what": f"sha1 of every rank's own columns (base, water, wall, light) after upload + {K} iterations "
                                                               "vs the same run on one GPU (rank 0's device)"}
            del full
    elif not args.no_prewarm:  # (profiler runs skip the long leg)
        e2e_long_s = e2e_run(E2E_LONG_K)
        e2e["k200"] = {"value": W * H * E2E_LONG_K / e2e_long_s, "steps": E2E_LONG_K, "seconds": e2e_long_s,
                       "h2d_bytes_per_step": h2d_total / E2E_LONG_K + 56, "d2h_bytes_per_step": d2h_total / E2E_LONG_K}
    sim.close()
    del hb, hw, hl, ob, ow, ol

    partition = "single GPU"
    if world > 1:
        how = {"peer": "k_push_ghosts: edge columns stored straight into the neighbours' ghost columns over NVLink (cudaIpc windows), sequence flags",
               "peerc": "k_push_ghosts: edge columns stored into the neighbours' compact landing zones over NVLink (cudaIpc windows), sequence flags, "
                        "k_unpack_zone copies them into the ghost columns"}.get(transport, "pack + ncclSend/ncclRecv + unpack")
        partition = f"{world} x-strip(s) of {lw} columns, ghost {gh}, one ring exchange per iteration ({how})"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full physics {W}x{H} fp32 (pressure+velocity+vorticity+boundary+advection+condensation+lighting), no particles",
                       "grid": [W, H], "partition": partition, "transport": transport if world > 1 else None,
                       "transport_calibration_ms_per_step": calibration,
                       "schedule": "fused: k_fused_pvb + k_fused_adv per iteration (TMA-staged channel planes)", "prewarm": f"{PREWARM_ITERS} untimed iterations before the W warm-up steps (SM clock ramp-up after host-side state generation)", "l2": "no flush: every plane is >= 256 MiB, far larger than the 126 MB L2",
                       "max_abs_velocity_cells_per_iter": vmax, "state_finite": finite},
            "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if ident is not None:
        line["config"]["bit_identical_to_1gpu"] = ident["ok"]
        line["config"]["bit_identical_check"] = ident

    if world == 1 and rank == 0 and not args.headline_only:
        pre = not args.no_prewarm
        line["dry_sweep"] = dry_sweep_leg(W, H, K, Wm, peak, peak_src, local_rank, prewarm=pre)
        line["config2_dry_4096x1024"] = dry_sweep_leg(4096, 1024, K, Wm, peak, peak_src, local_rank, prewarm=pre,
                                                      note="working set (4 base planes x 2 copies + wall = 151 MB) is L2-sized (126 MB L2): the HBM roofline fraction is not meaningful here (SURVEY 8d)")
        line["config3_full_8192x2048"] = full_leg(8192, 2048, K, Wm, peak, peak_src, local_rank, prewarm=pre)
        line["with_particles"] = particles_leg(W, H, K, Wm, peak, peak_src, local_rank, 1_000_000, prewarm=pre)
        if pre:
            line["with_particles_ref_count"] = particles_leg(W, H, K, Wm, peak, peak_src, local_rank, W * H // 25)
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
        if ident is not None and not ident["ok"]:
            raise SystemExit("bench.py: strips differ from the single-GPU run: " + ", ".join(ident["mismatches"]))


def dry_sweep_leg(W, H, K, Wm, peak, peak_src, device_index, prewarm=True, note=None):
    """The fused pressure+velocity+advection sweep (k_fused_dry): at the headline grid the kernel the
    >= 70 % HBM-roofline target of BASELINE.json is stated on; at 4096 x 1024 BASELINE config 2."""
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
    batches = _timed_batches(sim, sim.step_dry, K, SUSTAINED_BATCHES if prewarm else 1, 1, None)
    out = {"value": W * H * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "grid": [W, H],
           "workload": "dry sweep: pressure(prev) + velocity + semi-Lagrangian advection of the base field, one kernel per iteration",
           "roofline": _roofline("k_fused_dry", per, W * H, peak, peak_src, W, H, 1),
           "sustained_ms_per_step_median": float(np.median(batches)),
           "max_abs_velocity_cells_per_iter": sim.max_velocity, "clocks": clocks}
    if note:
        out["note"] = note
    sim.close()
    return out


def full_leg(W, H, K, Wm, peak, peak_src, device_index, prewarm=True):
    """BASELINE config 3: full physics (no particles) at 8192 x 2048 on one GPU."""
    import wsb200

    S, P = wsb200.sim, wsb200.params
    g = P.resolve_settings(None)
    g["enablePrecipitation"] = False
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    sim = wsb200.Simulation(W, H, 0, device=device_index, gui_controls=g)
    base, water, wall, _ = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=False)
    sim.upload(base, water, wall)
    del base, water, wall
    sim.set_profiling(True)
    if prewarm:
        sim.step(4 * PREWARM_ITERS)
        sim.sync()
    sim.step(Wm)
    sim.sync()
    sampler = ClockSampler(device_index)
    sampler.start()
    sim.step(K)
    sim.sync()
    clocks = sampler.result()
    ms = sim.last_step_ms()
    kt = {n: sim.kernel_time_ms(k) for n, k in (("k_fused_pvb", S.KERNEL_PVB), ("k_fused_adv", S.KERNEL_ADV))}
    dom = max(kt, key=lambda n: kt[n][0])
    batches = _timed_batches(sim, sim.step, K, SUSTAINED_BATCHES if prewarm else 1, 1, None)
    rl = _roofline(dom, kt[dom][0] / max(kt[dom][1], 1), W * H, peak, peak_src, W, H, 1)
    rl["kernels_ms_per_step"] = {n: t / max(c, 1) for n, (t, c) in kt.items()}
    rl["step_frac_of_104B_roofline"] = (B_ALG_STEP_FULL * W * H / (ms / K * 1e-3) / 1e9) / peak
    out = {"value": W * H * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "grid": [W, H],
           "workload": "full physics (pressure+velocity+vorticity+boundary+advection+condensation+lighting), no particles",
           "roofline": rl, "sustained_ms_per_step_median": float(np.median(batches)), "clocks": clocks,
           "note": "every plane is 64 MiB: the 25 planes of one iteration (1.6 GB) are far beyond the 126 MB L2"}
    sim.close()
    return out


def particles_leg(W, H, K, Wm, peak, peak_src, device_index, nd, prewarm=True):
    """BASELINE config 4: full physics + precipitation particles on one GPU (1 M droplets as named by the config; and the
    W*H/25 droplets the reference itself allocates for this grid, app.js:452,1282)."""
    import wsb200

    S, P = wsb200.sim, wsb200.params
    g = P.resolve_settings(None)
    g["dayNightCycle"] = False
    g["sunAngle"] = 60.0
    sim = wsb200.Simulation(W, H, nd, device=device_index, gui_controls=g)
    base, water, wall, drops = wsb200.synth.full_state(W, H, seed=7, g=g, with_droplets=True, n_droplets=nd)
    wsb200.synth.add_clouds(base, water, wall, n_blobs=96, seed=5)  # something to rain from
    sim.upload(base, water, wall, drops)
    del base, water, wall
    sim.set_profiling(True)
    sim.step(max(Wm, 30) + (PREWARM_ITERS if prewarm else 0))  # spin-up: the first iterations spawn the bulk of the droplets
    sim.sync()
    sampler = ClockSampler(device_index)
    sampler.start()
    sim.step(K)
    sim.sync()
    clocks = sampler.result()
    ms = sim.last_step_ms()
    kt = {n: sim.kernel_time_ms(k) for n, k in (("k_fused_pvb", S.KERNEL_PVB), ("k_fused_adv", S.KERNEL_ADV), ("particle_pass", S.KERNEL_PRECIP),
                                                ("of_which_boxsum_and_clear", S.KERNEL_SPRITES))}
    d = sim.read_droplets()
    dom = max(("k_fused_pvb", "k_fused_adv"), key=lambda n: kt[n][0])
    rl = _roofline(dom, kt[dom][0] / max(kt[dom][1], 1), W * H, peak, peak_src, W, H, 1)
    if dom == "k_fused_pvb":  # after a particle pass the boundary kernel also reads feedback (16 B) + deposition (8 B)
        rl["alg_bytes_per_cell"] = B_ALG[dom] + 24
        rl["achieved"] = rl["alg_bytes_per_cell"] * W * H / (rl["avg_launch_ms"] * 1e-3) / 1e9
        rl["frac"] = rl["achieved"] / peak
    rl["traffic"], rl["traffic_source"] = None, None
    rl["kernels_ms_per_step"] = {n: t / max(c, 1) for n, (t, c) in kt.items()}
    out = {"value": W * H * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "grid": [W, H], "n_droplets": nd,
           "active_droplets": int((d[:, 2] >= 0).sum()), "workload": "full physics + precipitation particles (k_precipitation + k_latch per iteration)",
           "roofline": rl, "clocks": clocks}
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
    ap.add_argument("--headline-only", action="store_true", help="N=1: skip the extra legs (dry sweep, configs 2-4, cpu baseline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="experiments: headline batch + 3 sustained batches only (no e2e, no 1-GPU comparison)")
    ap.add_argument("--no-prewarm", action="store_true", help="skip the clock ramp-up iterations (profiler runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
