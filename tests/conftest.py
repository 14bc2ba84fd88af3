import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
