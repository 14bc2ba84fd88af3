"""The C-ABI library: loads without a GPU, exports every symbol include/wsb200.h declares, and its
struct layouts match the ctypes mirrors.  No compute calls here."""
import ctypes
import os
import re

import pytest

import wsb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "wsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(wsb_[a-z_0-9]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_what_python_binds():
    assert _declared_functions() == sorted(wsb200.sim.EXPORTS)


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library)
    for name in _declared_functions():
        assert hasattr(lib, name), f"libwsb200.so does not export {name}"
    lib.wsb_build_info.restype = ctypes.c_char_p
    info = lib.wsb_build_info().decode()
    assert "sm_100a" in info


def test_struct_layouts_match_header():
    # wsb_params: 30 floats + 2 int32; wsb_frame_inputs: 2 + 4 + 2 floats, 2 int32, 4 floats
    assert ctypes.sizeof(wsb200.params.WsbParams) == 128
    assert ctypes.sizeof(wsb200.params.WsbFrameInputs) == 56
    assert ctypes.sizeof(wsb200.sim.WsbConfig) == 8 * 4 + 128
    text = open(os.path.join(ROOT, "include", "wsb200.h")).read()
    body = text[text.index("typedef struct wsb_params {"):text.index("} wsb_params;")]
    header_fields = re.findall(r"^\s*(?:float|int32_t)\s+([A-Za-z_0-9]+);", body, flags=re.M)
    assert header_fields == [n for n, _ in wsb200.params.WsbParams._fields_]


def test_error_paths_without_gpu(built_library):
    """Argument validation happens before any CUDA call, so it is testable on a CPU box; the
    device check itself must fail loudly (no CPU fallback)."""
    L = wsb200.load_library()
    cfg = wsb200.sim.WsbConfig()
    h = ctypes.c_void_p()
    cfg.abi_version = 999
    assert L.wsb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"abi_version" in L.wsb_last_error()
    cfg.abi_version = wsb200.sim.ABI_VERSION
    cfg.width, cfg.height = 8, 8
    assert L.wsb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"outside" in L.wsb_last_error()
    cfg.width, cfg.height, cfg.n_ranks, cfg.rank = 128, 128, 2, 5
    assert L.wsb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert L.wsb_step(None, 1) != 0
    assert L.wsb_read_rect(None, 0, 0, 0, 0, 1, 1, None) != 0


def test_no_cpu_fallback_when_no_device(built_library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wsb200.sim.WsbError):
        wsb200.Simulation(128, 128, 0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "2d-weather-sandbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
