"""Worker of tests/test_gpu_strips.py (not a test module): x-strip runs whose ranks all live on ONE
GPU, so that the strip decomposition, the peer-memory ghost exchange (k_push_ghosts / k_wait_ghosts,
csrc/wsb200.cu) and its guards are checked on a single-GPU box.

    python tests/strips_worker.py inproc  N W H iters [dry]   N sims in this process, plain pointers
    python tests/strips_worker.py ipc     N W H iters         N processes on device 0, cudaIpc handles
    python tests/strips_worker.py guard   N W H               |v| beyond the ghost budget must fail
    python tests/strips_worker.py timeout W H                 a neighbour that never steps must fail

Every mode compares with a single-GPU run of the same state (bit for bit) and exits non-zero on a
mismatch.  The parent sets CUDA_DEVICE_MAX_CONNECTIONS so that every stream of every rank gets its
own hardware queue (a spinning wait kernel must never sit in front of a neighbour's work)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FIELDS = (("base", 0, 0), ("water", 1, 1), ("wall", 2, 0), ("light", 3, 2))  # name, field id, view


def state(W, H):
    from util import stress_state

    g, base, water, wall, _ = stress_state(W, H, seed=23)
    g["enablePrecipitation"] = False
    return g, base, water, wall


def single_gpu(W, H, iters, dry=False):
    import wsb200
    from util import make_cuda

    g, base, water, wall = state(W, H)
    one = make_cuda(g, base, water, wall, None, wsb200.sim.SCHEDULE_FUSED)
    (one.step_dry if dry else one.step)(iters)
    out = {name: one.read_pixels(f, view=v) for name, f, v in FIELDS}
    vmax = one.max_velocity
    one.close()
    return out, vmax


def make_ring(n, W, H, g):
    import wsb200

    sims = [wsb200.Simulation(W, H, 0, device=0, rank=r, n_ranks=n, comm_id=None, gui_controls=g) for r in range(n)]
    infos = [s.peer_info() for s in sims]
    for r, s in enumerate(sims):
        s.connect_peers(infos[(r - 1) % n], infos[(r + 1) % n])
        if os.environ.get("WSB_TEST_EXCHANGE"):  # "peerc": the landing-zone variant of the peer transport
            s.set_exchange(os.environ["WSB_TEST_EXCHANGE"])
    return sims


def compare(got, want, what):
    bad = [name for name, _, _ in FIELDS if not np.array_equal(got[name], want[name])]
    if bad:
        for name in bad:
            d = np.argwhere((got[name] != want[name]).any(axis=-1))
            print(f"{what}: {name} differs in {len(d)} cells, first (y, x) = {d[0]}, columns {sorted(set(d[:, 1]))[:12]}")
        sys.exit(1)


def inproc(n, W, H, iters, dry):
    import wsb200

    g, base, water, wall = state(W, H)
    sims = make_ring(n, W, H, g)
    for s in sims:
        s.upload(base, water, wall, None)
        s.set_frame_inputs(wsb200.params.frame_inputs(g))
    done = 0
    while done < iters:  # small batches: every rank's work is enqueued before any wait can grow old
        k = min(3, iters - done)
        for s in sims:
            (s.step_dry if dry else s.step)(k)
        done += k
    got = {name: np.zeros((H, W, 4), np.int8 if name == "wall" else np.float32) for name, _, _ in FIELDS}
    for s in sims:
        for name, f, v in FIELDS:
            s.read_pixels(f, view=v, out=got[name])  # fills only the rank's own columns
    want, vmax = single_gpu(W, H, iters, dry)
    if dry:
        got, want = {"base": got["base"]}, {"base": want["base"]}
        bad = not np.array_equal(got["base"], want["base"])
        if bad:
            print("dry strips differ from the single-GPU run")
            sys.exit(1)
    else:
        compare(got, want, f"{n} in-process strips")
    launches = sims[0].launch_count
    for s in sims:
        s.close()
    print(f"OK inproc n={n} {W}x{H} iters={iters} dry={dry} vmax={vmax:.3f} launches(rank0)={launches}")


def _ipc_rank(rank, n, port, W, H, iters, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import wsb200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=n)
    try:
        g, base, water, wall = state(W, H)
        sim = wsb200.multi.create_distributed(W, H, device=0, gui_controls=g, transport=os.environ.get("WSB_TEST_EXCHANGE", "peer"))  # every rank on GPU 0
        sim.upload(base, water, wall, None)
        sim.set_frame_inputs(wsb200.params.frame_inputs(g))
        sim.step(iters)
        x0, lw = sim.strip()
        out = {}
        for name, f, v in FIELDS:
            full = sim.read_pixels(f, view=v)
            out[name] = wsb200.multi.gather_strips(full[:, x0:x0 + lw], W)
        if rank == 0:
            np.savez(ret, **out)
        dist.barrier()  # nobody unmaps a window a neighbour may still be writing into
        sim.close()
    finally:
        dist.destroy_process_group()


def ipc(n, W, H, iters):
    import tempfile

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = os.path.join(tempfile.mkdtemp(), "out.npz")
    mp.spawn(_ipc_rank, args=(n, port, W, H, iters, ret), nprocs=n, join=True)
    got = np.load(ret)
    want, vmax = single_gpu(W, H, iters)
    compare(got, want, f"{n} processes on one GPU (cudaIpc)")
    print(f"OK ipc n={n} {W}x{H} iters={iters} vmax={vmax:.3f}")


def guard(n, W, H):
    """A flow faster than the ghost budget: the strips must FAIL at the next synchronisation."""
    import wsb200

    g, base, water, wall = state(W, H)
    base[..., 0] = np.where(wall[..., 1] != 0, np.float32(6.5), base[..., 0])  # 6.5 cells / iteration everywhere in the air
    sims = make_ring(n, W, H, g)
    for s in sims:
        s.upload(base, water, wall, None)
    for s in sims:
        s.step(2)
    failed = 0
    for s in sims:
        try:
            s.sync()
        except wsb200.sim.WsbError as e:
            failed += 1
            msg = str(e)
    for s in sims:
        s.close()
    if failed != n or "ghost-zone budget" not in msg:
        print(f"guard: only {failed} of {n} strips refused a flow beyond the ghost budget")
        sys.exit(1)
    print(f"OK guard n={n}: {msg}")


def timeout(W, H):
    """Rank 1 never steps: rank 0's bounded waits must run out and surface as an error, not a hang."""
    import wsb200

    g, base, water, wall = state(W, H)
    sims = make_ring(2, W, H, g)
    for s in sims:
        s.upload(base, water, wall, None)
    sims[0].step(2)
    try:
        sims[0].sync()
    except wsb200.sim.WsbError as e:
        print(f"OK timeout: {e}")
        for s in sims:
            s.close()
        return
    print("timeout: a strip whose neighbour never stepped synchronised without an error")
    sys.exit(1)


if __name__ == "__main__":
    mode = sys.argv[1]
    a = [int(v) for v in sys.argv[2:6] if v.isdigit()]
    if mode == "inproc":
        inproc(a[0], a[1], a[2], a[3], "dry" in sys.argv)
    elif mode == "ipc":
        ipc(a[0], a[1], a[2], a[3])
    elif mode == "guard":
        guard(a[0], a[1], a[2])
    elif mode == "timeout":
        timeout(a[0], a[1])
    else:
        raise SystemExit(f"unknown mode {mode}")
