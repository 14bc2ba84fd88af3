"""2/4/8-GPU x-strip runs vs the single-GPU run: bit-identical fields, both ghost-exchange transports
(needs >= 2 CUDA devices; on a 1-GPU box these are skipped — there tests/test_gpu_strips.py runs the
same strips with every rank on GPU 0, and the decomposition plan itself is covered on CPU by
test_oracle_sim.py::test_fake_cluster_strips_bit_identical and test_multi_gloo.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, iters, ret, transport):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import wsb200
    from util import stress_state

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g, base, water, wall, _ = stress_state(W, H, seed=23)
        g["enablePrecipitation"] = False
        sim = wsb200.multi.create_distributed(W, H, device=rank, gui_controls=g, transport=transport)
        sim.upload(base, water, wall, None)
        if transport == "auto":  # both transports set up; switching between them mid-run must not change a bit
            sim.step(iters // 2)
            timings = wsb200.multi.calibrate_exchange(sim, iters=2)
            assert timings["chosen"] in ("peer", "nccl") and sim.transport == timings["chosen"]
            sim.set_exchange("nccl" if timings["chosen"] == "peer" else "peer")
            sim.step(iters - iters // 2 - 8)   # calibrate_exchange advanced 4 * 2 iterations
        else:
            sim.step(iters)
        S = wsb200.sim
        x0, lw = sim.strip()
        out = {}
        for name, f, v in (("base", S.FIELD_BASE, 0), ("water", S.FIELD_WATER, 1), ("wall", S.FIELD_WALL, 0), ("light", S.FIELD_LIGHT, 2)):
            full = sim.read_pixels(f, view=v)
            out[name] = wsb200.multi.gather_strips(full[:, x0:x0 + lw], W)
        if rank == 0:
            np.savez(ret, **out)
        dist.barrier()  # nobody unmaps a window a neighbour may still be writing into
        sim.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "peerc", "nccl", "auto"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_strips_bit_identical_to_single_gpu(world, transport, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import wsb200
    from util import make_cuda, stress_state

    W, H, iters = 512, 96, 20
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, port, W, H, iters, ret, transport), nprocs=world, join=True)
    got = np.load(ret)
    g, base, water, wall, _ = stress_state(W, H, seed=23)
    g["enablePrecipitation"] = False
    S = wsb200.sim
    one = make_cuda(g, base, water, wall, None, S.SCHEDULE_FUSED)
    one.step(iters)
    assert np.array_equal(got["wall"], one.read_pixels(S.FIELD_WALL))
    assert np.array_equal(got["base"], one.read_pixels(S.FIELD_BASE))
    assert np.array_equal(got["water"], one.read_pixels(S.FIELD_WATER, view=1))
    assert np.array_equal(got["light"], one.read_pixels(S.FIELD_LIGHT, view=2))
    one.close()
