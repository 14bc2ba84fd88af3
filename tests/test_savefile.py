"""`.weathersandbox` codec (loadData app.js:1256-1366, prepareDownload app.js:6575-6628)."""
import json
import struct
import zlib

import numpy as np
import pytest

import wsb200

S = wsb200.savefile


def test_golden_save_parses(save100):
    sf = save100
    assert (sf.width, sf.height) == (100, 100)
    assert sf.version == S.SAVE_FILE_VERSION_ID
    assert sf.base.shape == (100, 100, 4) and sf.water.shape == (100, 100, 4) and sf.wall.shape == (100, 100, 4)
    assert sf.droplets.shape == (400, 5)  # W*H/25
    assert sf.stations.shape == (0, 2)
    assert not np.isnan(sf.base).any() and not np.isnan(sf.water).any()
    assert (sf.wall[0, :, 1] == 0).all()  # bottom row is wall
    assert set(np.unique(sf.wall[..., 0])) <= {1, 2}
    settings = json.loads(sf.settings_json)
    assert len(settings) == 51
    assert settings["evapHeat"] == 1.9
    assert (sf.droplets[:, 2] >= 0).sum() == 1  # one active droplet


def test_payload_round_trip_is_byte_identical(save100):
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "100x100_test.weathersandbox")
    blob = open(path, "rb").read()
    original_payload = zlib.decompress(blob[4:])
    assert S.payload(save100) == original_payload
    again = S.loads(S.dumps(save100))
    assert S.payload(again) == original_payload
    assert struct.unpack_from("<I", S.dumps(save100), 0)[0] == 263574036


def test_legacy_and_bad_versions():
    sf = S.SaveFile(32, 32, np.zeros((32, 32, 4), np.float32), np.zeros((32, 32, 4), np.float32),
                    np.zeros((32, 32, 4), np.int8), np.zeros((S.num_droplets(32, 32), 5), np.float32), version=S.LEGACY_VERSION_ID)
    back = S.loads(S.dumps(sf))
    assert back.settings_json is None and back.version == S.LEGACY_VERSION_ID
    with pytest.raises(S.IncompatibleFile):
        S.loads(struct.pack("<I", 42) + zlib.compress(b"xx"))
    with pytest.raises(S.IncompatibleFile):
        S.loads(b"\x01")
    good = S.dumps(sf)
    trunc = struct.pack("<I", S.LEGACY_VERSION_ID) + zlib.compress(zlib.decompress(good[4:])[:1000])
    with pytest.raises(S.IncompatibleFile):
        S.loads(trunc)


def test_ragged_droplet_count():
    # W*H not divisible by 25: the typed-array slicing floors
    assert S.num_droplets(33, 32) == (33 * 32) // 25
    sf = S.SaveFile(33, 32, np.ones((32, 33, 4), np.float32), np.ones((32, 33, 4), np.float32), np.ones((32, 33, 4), np.int8),
                    np.ones((S.num_droplets(33, 32), 5), np.float32), np.array([[3, 4]], np.int16), '{"a":1}')
    back = S.loads(S.dumps(sf))
    assert back.width == 33 and back.height == 32
    assert np.array_equal(back.stations, [[3, 4]])
    assert back.settings_json == '{"a":1}'


def test_native_codec_is_a_standard_zlib_stream(built_library):
    """libwsbsave.so: chunked multi-threaded deflate stitched into one zlib stream."""
    import time

    L = S.native_codec()
    assert L is not None, "libwsbsave.so not built (make -C 2d-weather-sandbox_b200/csrc)"
    rng = np.random.default_rng(0)
    for n in (0, 1, 1000, (4 << 20) - 1, 4 << 20, (4 << 20) + 1, 13_000_000):
        raw = (rng.integers(0, 8, n, dtype=np.uint8) * 17).tobytes()  # compressible, not trivial
        for threads in (1, 4):
            z = S.deflate(raw, 6, threads)
            assert zlib.decompress(z) == raw          # any inflate (pako, zlib) reads it
            assert S.inflate(z) == raw
    # and the other direction: a stream written by plain zlib
    raw = bytes(range(256)) * 5000
    assert S.inflate(zlib.compress(raw, 9)) == raw
    with pytest.raises(zlib.error):
        S.inflate(b"\x78\x9c" + b"garbage-garbage-garbage")
    # parallel deflate is faster than one thread on a field-sized buffer
    big = np.sin(np.arange(6_000_000, dtype=np.float32)).tobytes()
    t0 = time.perf_counter(); a = S.deflate(big, 6, 1); t1 = time.perf_counter(); b = S.deflate(big, 6, 4); t2 = time.perf_counter()
    assert zlib.decompress(a) == big and zlib.decompress(b) == big
    assert (t2 - t1) < (t1 - t0)


# ---------------------------------------------------------------------------------------------
# every save the reference ships (when the checkout is present), plus the committed fixtures: the C++ reader / writer
# (csrc/wsb_save.cpp) against the Python restatement and against the file's own bytes
# ---------------------------------------------------------------------------------------------
from test_reference_saves import SAVES  # noqa: E402


@pytest.mark.parametrize("name,path", [(s_[0], s_[1]) for s_ in SAVES], ids=[s_[0] for s_ in SAVES])
def test_native_container_round_trips_every_save(name, path, built_library):
    assert S.native_codec() is not None, "libwsbsave.so must be built (make -C 2d-weather-sandbox_b200/csrc)"
    blob = open(path, "rb").read()
    original = zlib.decompress(blob[4:])                      # what pako.inflate hands to loadData
    sf = S.loads(blob)                                         # C++ inflate + wsb_save_parse
    assert sf.droplets.shape == (sf.width * sf.height // 25, 5)
    assert S.payload(sf) == original                           # wsb_save_serialise reassembles the reference's bytes
    assert S.payload(sf, native=False) == original             # ... and so does the Python writer
    assert zlib.decompress(S.dumps(sf, level=1)[4:]) == original   # multi-threaded deflate -> one standard zlib stream
    if sf.settings_json is not None:
        json.loads(sf.settings_json)


def test_native_parser_rejects_what_loaddata_rejects(built_library):
    import ctypes

    L = S.native_codec()
    lay = S.SaveLayout()
    good = np.frombuffer(S.payload(S.SaveFile(32, 32, np.zeros((32, 32, 4), np.float32), np.zeros((32, 32, 4), np.float32), np.zeros((32, 32, 4), np.int8),
                                              np.zeros((S.num_droplets(32, 32), 5), np.float32), settings_json="{}")), np.uint8)
    assert L.wsb_save_parse(good.ctypes.data, good.size, S.SAVE_FILE_VERSION_ID, ctypes.byref(lay)) == 0
    assert (lay.width, lay.height, lay.n_droplets, lay.n_stations, lay.settings_len) == (32, 32, 40, 0, 2)
    assert L.wsb_save_parse(good.ctypes.data, good.size, 42, ctypes.byref(lay)) == -4            # unknown version id
    assert L.wsb_save_parse(good.ctypes.data, good.size - 10, S.SAVE_FILE_VERSION_ID, ctypes.byref(lay)) == -5   # truncated
    assert L.wsb_save_parse(good.ctypes.data, 3, S.SAVE_FILE_VERSION_ID, ctypes.byref(lay)) == -5
