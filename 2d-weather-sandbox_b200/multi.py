"""Multi-GPU host plumbing: one process per GPU (torchrun).  torch.distributed only carries the
bootstrap blobs (peer-memory windows, or the NCCL id) and the readback gathers; the per-iteration
ghost exchange itself runs inside libwsb200.so (csrc/wsb200.cu: exchange()): by default one kernel
that stores the edge columns straight into the neighbours' ghost columns over NVLink
(transport "peer"), optionally ncclSend / ncclRecv through staging buffers (transport "nccl").

The reference is single-GPU; this is the x-strip decomposition of SURVEY 8e.
"""
from __future__ import annotations

import numpy as np

from . import strips

COMM_ID_BYTES = 128
PEERC_MIN_STRIP = 4096  # strip width from which "auto" also times the landing-zone push (profiles/r3_multi_gpu.md)


def broadcast_comm_id(create_fn, group=None) -> bytes:
    """Rank 0 calls create_fn() -> 128-byte NCCL unique id; every rank returns the same bytes."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    box = [create_fn() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    cid = box[0]
    if not isinstance(cid, (bytes, bytearray)) or len(cid) != COMM_ID_BYTES:
        raise ValueError("comm id must be 128 bytes")
    return bytes(cid)


def connect_ring(sim, group=None) -> None:
    """All-gather every rank's peer window and link `sim` with its two ring neighbours."""
    import torch.distributed as dist

    n, r = dist.get_world_size(group), dist.get_rank(group)
    infos = [None] * n
    dist.all_gather_object(infos, sim.peer_info(), group=group)
    left, right = strips.neighbours(r, n)
    sim.connect_peers(infos[left], infos[right])


def calibrate_exchange(sim, iters: int = 12, group=None) -> dict:
    """Time `iters` iterations of the LIVE state with each ghost-exchange transport of a strip that has both
    (transport "auto"), max over ranks, and keep the faster one.  Which one wins depends on the strip width — measured
    on one NVSwitch box: the peer push at 2 and 8 GPUs, NCCL send/recv at 4 (profiles/r3_multi_gpu.md) — and both
    give bit-identical fields, so this is a pure scheduling decision.  Rings with two distinct neighbours and strips
    of 4096 columns or more — where the direct push is slow because its stores are scattered over every page of two
    peer arenas — also try "peerc", the push through a compact landing zone.  Advances the simulation by
    2 * iters iterations per candidate; every rank must call it at the same point.  Returns the timings (ms per
    iteration)."""
    import torch
    import torch.distributed as dist

    names = tuple(getattr(sim, "transports", None) or ())
    if len(names) < 2:
        return {}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    out = {}
    for name in reversed(names):
        sim.set_exchange(name)
        sim.step(iters)  # settle into the transport's steady state
        sim.sync()
        dist.barrier(group)
        sim.step(iters)
        sim.sync()
        t = torch.tensor([sim.last_step_ms() / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        out[name] = float(t.item())
    best = min(out, key=out.get)
    sim.set_exchange(best)
    out["chosen"] = best
    return out


def create_distributed(width: int, height: int, *, device: int, gui_controls=None, group=None, transport: str | None = None, **kw):
    """A Simulation for this rank's strip of a width x height grid (torch.distributed must be
    initialised; with world size 1 this is a plain single-GPU simulation).  transport: "peer"
    (default; WSB_EXCHANGE overrides), "peerc" (peer through a compact landing zone), "nccl", or "auto" = all set
    up, peer selected until calibrate_exchange(sim) has timed them on the live state."""
    import os

    import torch.distributed as dist

    from .sim import Simulation, comm_id_create

    n = dist.get_world_size(group) if dist.is_initialized() else 1
    r = dist.get_rank(group) if dist.is_initialized() else 0
    if n == 1:
        return Simulation(width, height, kw.pop("n_droplets", 0), device=device, gui_controls=gui_controls, **kw)
    transport = transport or os.environ.get("WSB_EXCHANGE", "peer")
    if transport not in ("peer", "peerc", "nccl", "auto"):
        raise ValueError(f"unknown ghost-exchange transport {transport!r}")
    kw.pop("n_droplets", None)
    if transport == "nccl":
        cid = broadcast_comm_id(comm_id_create, group)
        sim = Simulation(width, height, 0, device=device, rank=r, n_ranks=n, comm_id=cid, gui_controls=gui_controls, **kw)
        sim.transport, sim.transports = "nccl", ("nccl",)
        return sim
    # peer transport; if any rank cannot map its neighbours (no peer access between the devices, IPC disabled in a
    # container) EVERY rank falls back to the NCCL transport — loudly
    cid = broadcast_comm_id(comm_id_create, group) if transport == "auto" else None
    sim, err = None, None
    try:
        sim = Simulation(width, height, 0, device=device, rank=r, n_ranks=n, comm_id=cid, gui_controls=gui_controls, **kw)
        connect_ring(sim, group)
    except Exception as e:  # noqa: BLE001 — collected and agreed on below
        err = repr(e)
    errs = [None] * n
    dist.all_gather_object(errs, err, group=group)
    if all(e is None for e in errs):
        sim.transport = "peer"
        if transport == "peerc":
            sim.set_exchange("peerc")
            sim.transports = ("peerc",)
        elif transport == "auto":
            wide_ring = n >= 3 and width // n >= PEERC_MIN_STRIP  # rank-independent: every rank must time the same candidates
            sim.transports = ("peer", "nccl", "peerc") if wide_ring else ("peer", "nccl")
        else:
            sim.transports = ("peer",)
        return sim
    if sim is not None:
        sim.close()
    if os.environ.get("WSB_EXCHANGE") == "peer":  # explicitly requested: do not hide the failure
        raise RuntimeError(f"peer-memory ghost exchange unavailable: {[e for e in errs if e][0]}")
    if r == 0:
        import warnings

        warnings.warn(f"wsb200: peer-memory ghost exchange unavailable ({[e for e in errs if e][0]}); using ncclSend/ncclRecv")
    cid = broadcast_comm_id(comm_id_create, group)
    sim = Simulation(width, height, 0, device=device, rank=r, n_ranks=n, comm_id=cid, gui_controls=gui_controls, **kw)
    sim.transport, sim.transports = "nccl", ("nccl",)
    return sim


def gather_strips(local: np.ndarray, width: int, group=None):
    """All-gather per-rank strips [H][local_width][C] into the global [H][W][C] array (every rank
    gets the result).  Strip r covers strips.strip_bounds(width, n, r)."""
    import torch
    import torch.distributed as dist

    n = dist.get_world_size(group)
    parts = [None] * n
    dist.all_gather_object(parts, np.ascontiguousarray(local), group=group)
    h, c = local.shape[0], local.shape[2]
    out = np.zeros((h, width, c), local.dtype)
    for r, part in enumerate(parts):
        x0, lw = strips.strip_bounds(width, n, r)
        if part.shape != (h, lw, c):
            raise ValueError(f"rank {r}: strip shape {part.shape} != {(h, lw, c)}")
        out[:, x0:x0 + lw] = part
    del torch
    return out


def ring_exchange(padded: list[np.ndarray], local_width: int, group=None) -> None:
    """Reference implementation of the ghost-column exchange plan on HOST arrays (what
    csrc/wsb200.cu: exchange() does with NCCL on device arrays): every array is
    [H][local_width + 2*GHOST][C]; after the call the ghost columns hold the neighbours' edge
    columns.  Used by the CPU (gloo) tests of the decomposition plan."""
    import torch
    import torch.distributed as dist

    n, r = dist.get_world_size(group), dist.get_rank(group)
    left, right = strips.neighbours(r, n)
    snd, rcv = strips.send_slices(local_width), strips.recv_slices(local_width)
    for a in padded:
        to_left = torch.from_numpy(np.ascontiguousarray(a[:, snd["to_left"]]))
        to_right = torch.from_numpy(np.ascontiguousarray(a[:, snd["to_right"]]))
        from_left = torch.empty_like(to_right)
        from_right = torch.empty_like(to_left)
        # receives posted (right, left) — same pairing rule as the NCCL path, see exchange()
        ops = [dist.P2POp(dist.isend, to_left, left, group, tag=1), dist.P2POp(dist.isend, to_right, right, group, tag=2),
               dist.P2POp(dist.irecv, from_right, right, group, tag=1), dist.P2POp(dist.irecv, from_left, left, group, tag=2)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        a[:, rcv["from_left"]] = from_left.numpy()
        a[:, rcv["from_right"]] = from_right.numpy()
