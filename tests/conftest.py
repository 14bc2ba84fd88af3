import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST_CELLS = os.path.join(ROOT, "tests", "host_cells")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    """Driver-level probe (no torch import, no context): 0 on a CPU box."""
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU box skips the gpu-marked tests instead of failing in wsb_create."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The oracle is test infrastructure: build it once per session."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


@pytest.fixture(scope="session")
def built_library():
    """libwsb200.so, cross-compiled for sm_100a (nvcc works without a GPU)."""
    import wsb200

    if not os.path.exists(wsb200.library_path()) or not os.path.exists(os.path.join(os.path.dirname(wsb200.library_path()), "libwsbsave.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "2d-weather-sandbox_b200", "csrc"), "-s"])
    return wsb200.library_path()


@pytest.fixture(scope="session")
def save100():
    import wsb200

    return wsb200.savefile.load(os.path.join(GOLDEN, "100x100_test.weathersandbox"))


# ---- test infrastructure built from the product's own sources (tests/host_cells/) ----------------
@pytest.fixture(scope="session")
def lib():
    subprocess.check_call(["make", "-C", HOST_CELLS, "-s"])
    L = ctypes.CDLL(os.path.join(HOST_CELLS, "libhostcells.so"))
    import wsb200

    P = wsb200.params
    vp = ctypes.c_void_p
    L.hc_create.restype = vp
    L.hc_create.argtypes = [ctypes.c_int, ctypes.c_int]
    L.hc_destroy.argtypes = [vp]
    L.hc_upload.argtypes = [vp, vp, vp, vp]
    L.hc_set_params.argtypes = [vp, ctypes.POINTER(P.WsbParams)]
    L.hc_set_frame_inputs.argtypes = [vp, ctypes.POINTER(P.WsbFrameInputs)]
    L.hc_set_profiles.argtypes = [vp, vp, vp, vp, vp]
    L.hc_set_iter.argtypes = [vp, ctypes.c_longlong]
    L.hc_set_feedback.argtypes = [vp, vp, vp]
    L.hc_run_pass.argtypes = [vp, ctypes.c_int]
    L.hc_read.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    return L



def _load_emu(name="libemufused.so", flags=None):
    cmd = ["make", "-C", HOST_CELLS, "-s", name] + ([f"EMU_FLAGS={flags}"] if flags else [])
    subprocess.check_call(cmd)
    L = ctypes.CDLL(os.path.join(HOST_CELLS, name))
    import wsb200

    P = wsb200.params
    vp = ctypes.c_void_p
    L.ef_create.restype = vp
    L.ef_create.argtypes = [ctypes.c_int, ctypes.c_int]
    L.ef_create_strip.restype = vp
    L.ef_create_strip.argtypes = [ctypes.c_int] * 5
    L.ef_exchange_planes.argtypes = [vp, ctypes.POINTER(vp)]
    L.ef_destroy.argtypes = [vp]
    L.ef_upload.argtypes = [vp, vp, vp, vp]
    L.ef_set_params.argtypes = [vp, ctypes.POINTER(P.WsbParams)]
    L.ef_set_frame_inputs.argtypes = [vp, ctypes.POINTER(P.WsbFrameInputs)]
    L.ef_set_profiles.argtypes = [vp, vp, vp, vp, vp]
    L.ef_set_iter.argtypes = [vp, ctypes.c_longlong]
    L.ef_uses_tma.argtypes = [vp]
    L.ef_step.argtypes = [vp, ctypes.c_int]
    L.ef_step_dry.argtypes = [vp, ctypes.c_int]
    L.ef_max_velocity.argtypes = [vp]
    L.ef_max_velocity.restype = ctypes.c_float
    L.ef_read.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    L.ef_upload_drops.argtypes = [vp, vp, ctypes.c_int]
    L.ef_read_drops.argtypes = [vp, vp]
    L.ef_read_feedback.argtypes = [vp, vp, vp]
    L.ef_get_latches.argtypes = [vp, vp, vp]
    L.ef_set_inactive.argtypes = [vp, ctypes.c_float]
    L.ef_origin_residue.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    return L


@pytest.fixture(scope="session")
def emu():
    return _load_emu()
