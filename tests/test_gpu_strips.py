"""x-strip runs with every rank on ONE GPU (tests/strips_worker.py): the decomposition, the
peer-memory ghost exchange and its guards are checked against the single-GPU run, bit for bit, on
the single-GPU box the driver runs `-m gpu` on.  (The reference is single-GPU, SURVEY 5.8: the
1-GPU run is the oracle of the strips; 2/4/8 real GPUs: tests/test_gpu_multi.py and bench.py's
config.bit_identical_to_1gpu.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "strips_worker.py")


def _run(*args, env=None, timeout=600):
    e = dict(os.environ)
    e["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"  # one hardware queue per stream: a waiting rank never blocks its neighbour's launches
    e.update(env or {})
    r = subprocess.run([sys.executable, WORKER, *map(str, args)], env=e, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f"strips_worker {args} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    assert "OK " in r.stdout, r.stdout
    return r.stdout


@pytest.mark.parametrize("n,w,h,iters", [(2, 512, 96, 20), (3, 600, 80, 14), (4, 1024, 64, 12), (2, 2304, 160, 9), (4, 320, 64, 10), (8, 512, 96, 8)])
def test_strips_in_one_process_bit_identical_to_single_gpu(n, w, h, iters):
    """N sims of one host process linked with plain device pointers: full physics; strip widths that are and are not
    multiples of the tile width (600 / 3 = 200), strips with interior tile columns (two-stream schedule) and strips too
    narrow to have any (320 / 4 = 80 and 512 / 8 = 64 columns: one launch per kernel, exchange after it)."""
    _run("inproc", n, w, h, iters)


@pytest.mark.parametrize("n,w,h,iters", [(2, 512, 96, 20), (3, 600, 80, 14), (4, 1024, 64, 12), (4, 320, 64, 10)])
def test_landing_zone_push_in_one_process_bit_identical_to_single_gpu(n, w, h, iters):
    """Transport "peerc": the neighbours store into this rank's compact landing zones, k_unpack_zone copies them into
    the ghost columns — with one neighbour on both sides (n = 2), with two distinct neighbours (n = 3, 4), with interior
    tile columns (two-stream schedule) and without (320 / 4 = 80 columns)."""
    _run("inproc", n, w, h, iters, env={"WSB_TEST_EXCHANGE": "peerc"})


def test_landing_zone_push_in_two_processes_over_cuda_ipc():
    _run("ipc", 2, 512, 96, 12, env={"WSB_TEST_EXCHANGE": "peerc"})


def test_dry_strips_in_one_process_bit_identical_to_single_gpu():
    _run("inproc", 2, 512, 96, 10, "dry")


def test_strips_in_two_processes_on_one_gpu_over_cuda_ipc():
    """The product plumbing (multi.create_distributed over a gloo group, cudaIpc handles) with both ranks on GPU 0."""
    _run("ipc", 2, 512, 96, 12)


def test_flow_beyond_the_ghost_budget_is_refused():
    _run("guard", 2, 512, 96)


def test_missing_neighbour_becomes_an_error_not_a_hang():
    _run("timeout", 512, 96, env={"WSB_SPIN_LIMIT_MS": "300"}, timeout=120)
