"""World-size-2 (and 4) runs of the multi-GPU host plumbing on CPU with the gloo backend: NCCL-id
broadcast, strip bookkeeping, readback gather, and the ring ghost exchange plan driving oracle
strips — which must reproduce the whole-domain oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, iters, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import wsb200
    from oracle import oracle as O
    from util import stress_state

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, G = wsb200.params, wsb200.strips.GHOST
        cid = wsb200.multi.broadcast_comm_id(lambda: bytes(range(128)))
        assert cid == bytes(range(128))
        g, base, water, wall, _ = stress_state(W, H, seed=21)
        cols = wsb200.strips.padded_columns(W, world, rank)
        x0, lw = wsb200.strips.strip_bounds(W, world, rank)
        o = O.OracleSim(lw + 2 * G, H, 0, global_width=W, x0=x0 - G)
        o.upload(base[:, cols], water[:, cols], wall[:, cols], None)
        o.set_params(P.derive_params(g))
        o.set_frame_inputs(P.frame_inputs(g))
        o.set_profiles(P.initial_T_profile(H, g))
        fields = [(O.FIELD_BASE, 0), (O.FIELD_BASE, 1), (O.FIELD_WATER, 0), (O.FIELD_WATER, 1), (O.FIELD_WALL, 0),
                  (O.FIELD_WALL, 1), (O.FIELD_LIGHT, 0), (O.FIELD_LIGHT, 1)]
        for _ in range(iters):
            o.step(1)
            wsb200.multi.ring_exchange([o.field(f, b, copy=False) for f, b in fields], lw)
        out = {}
        for name, (f, b) in {"base": (O.FIELD_BASE, 0), "water": (O.FIELD_WATER, 1), "wall": (O.FIELD_WALL, 0), "light": (O.FIELD_LIGHT, 0)}.items():
            out[name] = wsb200.multi.gather_strips(o.field(f, b)[:, G:G + lw], W)
        if rank == 0:
            np.savez(ret, **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ring_exchange_gloo(world, tmp_path):
    import socket

    import wsb200
    from oracle import oracle as O
    from util import make_oracle, stress_state

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    W, H, iters = 160, 40, 6
    ret = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, port, W, H, iters, ret), nprocs=world, join=True)
    got = np.load(ret)
    g, base, water, wall, _ = stress_state(W, H, seed=21)
    whole = make_oracle(g, base, water, wall, None)
    whole.step(iters)
    assert np.array_equal(got["base"], whole.field(O.FIELD_BASE, 0))
    assert np.array_equal(got["water"], whole.field(O.FIELD_WATER, 1))
    assert np.array_equal(got["wall"], whole.field(O.FIELD_WALL, 0))
    assert np.array_equal(got["light"], whole.field(O.FIELD_LIGHT, 0))
    del wsb200


def test_strip_plan():
    import wsb200

    S = wsb200.strips
    assert S.strip_bounds(16384, 8, 0) == (0, 2048) and S.strip_bounds(16384, 8, 7) == (14336, 2048)
    assert S.neighbours(0, 8) == (7, 1) and S.neighbours(7, 8) == (6, 0)
    assert S.neighbours(0, 2) == (1, 1)
    cols = S.padded_columns(64, 2, 0)
    assert cols[:S.GHOST] == list(range(56, 64)) and cols[S.GHOST] == 0 and cols[-1] == 32 + S.GHOST - 1
    with pytest.raises(ValueError):
        S.strip_bounds(64, 8, 0)  # 8-column strips are narrower than 2*GHOST
    assert S.padded_columns(100, 1, 0) == list(range(100))
    # uneven split covers every column exactly once
    covered = []
    for r in range(3):
        x0, lw = S.strip_bounds(100, 3, r)
        covered += list(range(x0, x0 + lw))
    assert covered == list(range(100))
