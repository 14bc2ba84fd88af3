"""The N-API shim (bindings/node/wsb200_napi.c) built into a real shared object, loaded and driven — without Node.

bindings/node/fake_napi_host.c implements the Node-API subset the shim uses and plays the JavaScript side: it dlopen()s
the addon (whose constructor registers the module, as under Node), then calls create -> upload -> setParams ->
setProfiles -> setFrameInputs -> step -> readRect -> readDroplets -> getInactiveDroplets -> getLightning -> destroy,
the calls that replace app.js:5149-5317 (allocation, upload), 3401-3443 (uniforms), 5830-6005 (the loop) and its
gl.readPixels / getBufferSubData sites.  The GPU test holds the result to the same run through the ctypes mirror."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import wsb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE = os.path.join(ROOT, "bindings", "node")
EXPORTS = ["create", "destroy", "upload", "setParams", "setProfiles", "setFrameInputs", "step", "readRect", "readDroplets",
           "getInactiveDroplets", "getLightning"]


@pytest.fixture(scope="module")
def addon(built_library):
    subprocess.check_call(["make", "-C", NODE, "-s"])
    return os.path.join(NODE, "fake_napi_host"), os.path.join(NODE, "wsb200_napi.so")


def test_addon_builds_loads_and_registers_its_exports(addon):
    host, so = addon
    out = subprocess.run([host, so, "--list"], capture_output=True, text=True, check=True).stdout.split()
    assert out[:2] == ["module", "wsb200:"] and out[2:] == EXPORTS
    # the napi_* entry points are left to the host process, everything else resolves (libwsb200.so via rpath)
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True, check=True).stdout
    assert "napi_module_register" in undefined and "wsb_create" in undefined


@pytest.mark.gpu
def test_addon_drives_the_simulation_like_the_ctypes_host(addon, save100, tmp_path):
    host, so = addon
    sf = save100
    g = wsb200.params.resolve_settings(sf.settings_json)
    p = wsb200.params.derive_params(g)
    fi = wsb200.params.frame_inputs(g)
    t0 = wsb200.params.initial_T_profile(sf.height, g)
    iters = 25
    fields = [n for n, _ in p._fields_]
    pf = np.array([getattr(p, n) for n in fields[:30]], np.float32)
    fi_words = np.zeros(14, np.float32)
    fi_words[0], fi_words[1] = fi.sunAngle, fi.sunIntensity
    fi_words[2:6], fi_words[6:8] = list(fi.userInputValues), list(fi.userInputMove)
    fi_words[8:10] = np.array([fi.userInputType, fi.wrapHorizontally], np.int32).view(np.float32)
    fi_words[10:14] = list(fi.airplaneValues)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([sf.width, sf.height, sf.droplets.shape[0], iters], np.int32).tobytes())
        for a in (sf.base, sf.water, sf.wall, sf.droplets, pf, np.array([p.enablePrecipitation], np.int32), t0.astype(np.float32), fi_words):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([host, so, str(inp), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n = sf.width * sf.height
    raw = open(outp, "rb").read()
    got_base = np.frombuffer(raw, np.float32, n * 4, 0).reshape(sf.height, sf.width, 4)
    got_wall = np.frombuffer(raw, np.int8, n * 4, n * 16).reshape(sf.height, sf.width, 4)
    nd = sf.droplets.shape[0]
    got_drops = np.frombuffer(raw, np.float32, nd * 5, n * 20).reshape(nd, 5)
    tail = np.frombuffer(raw, np.float32, 5, n * 20 + nd * 20)
    sim = wsb200.Simulation.from_save(sf)
    sim.set_frame_inputs(fi)
    sim.step(iters)
    S = wsb200.sim
    assert np.array_equal(got_wall, sim.read_pixels(S.FIELD_WALL))
    assert np.array_equal(got_base, sim.read_pixels(S.FIELD_BASE))
    assert np.array_equal(got_drops, sim.read_droplets())
    assert tail[0] == sim.inactive_droplets and np.array_equal(tail[1:], sim.lightning)
    assert ctypes.sizeof(p) == 128  # 30 floats + enablePrecipitation + reserved: what setParams copies
    sim.close()
