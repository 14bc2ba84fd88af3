"""`.weathersandbox` save-file codec.

Restates loadData (reference app.js:1256-1366) and prepareDownload (app.js:6575-6628):

    u32 LE  version id  (263574036; 1939327491 = previous version without settings, app.js:1265)
    zlib(deflate) of:
        u16 W, u16 H
        f32 base [H][W][4]
        f32 water[H][W][4]
        i8  wall [H][W][4]
        f32 droplets[N][5]      N = W*H/25   (app.js:1282, NUM_DROPLETS_DEVIDER = 25)
        i16 nStations, i16 xy[nStations][2]          (current version only)
        UTF-8 JSON of guiControls, to end of stream  (current version only)

Rows are bottom-up (GL origin); everything is little endian.
"""
from __future__ import annotations

import ctypes
import os
import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CODEC = None


class SaveLayout(ctypes.Structure):
    """wsb_save_layout (include/wsb_save.h): byte offsets of the pieces of an inflated payload."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32)] + [
        (n, ctypes.c_int64) for n in ("n_droplets", "off_base", "off_water", "off_wall", "off_droplets", "n_stations", "off_stations",
                                      "off_settings", "settings_len")]


def native_codec():
    """libwsbsave.so (csrc/wsb_save.cpp): multi-threaded zlib codec, or None when it is not built
    (the pure-Python zlib path below produces the same container, on one thread)."""
    global _CODEC
    if _CODEC is None:
        path = os.path.join(_HERE, "csrc", "libwsbsave.so")
        if not os.path.exists(path):
            _CODEC = False
        else:
            L = ctypes.CDLL(path)
            i64, vp = ctypes.c_int64, ctypes.c_void_p
            L.wsb_save_compress_bound.restype = i64
            L.wsb_save_compress_bound.argtypes = [i64]
            L.wsb_save_compress.restype = i64
            L.wsb_save_compress.argtypes = [vp, i64, vp, i64, ctypes.c_int32, ctypes.c_int32]
            L.wsb_save_decompress.restype = i64
            L.wsb_save_decompress.argtypes = [vp, i64, vp, i64]
            L.wsb_save_inflated_size.restype = i64
            L.wsb_save_inflated_size.argtypes = [vp, i64]
            L.wsb_save_parse.restype = ctypes.c_int32
            L.wsb_save_parse.argtypes = [vp, i64, ctypes.c_uint32, ctypes.POINTER(SaveLayout)]
            L.wsb_save_payload_size.restype = i64
            L.wsb_save_payload_size.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, i64, ctypes.c_uint32]
            L.wsb_save_serialise.restype = i64
            L.wsb_save_serialise.argtypes = [vp, i64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp, vp, ctypes.c_int32, ctypes.c_char_p, i64,
                                             ctypes.c_uint32]
            _CODEC = L
    return _CODEC or None


def deflate(payload: bytes, level: int = 6, threads: int = 0) -> bytes:
    """zlib stream of `payload` (what pako.deflate produces in prepareDownload, app.js:6617)."""
    L = native_codec()
    if L is None:
        return zlib.compress(payload, level)
    src = np.frombuffer(payload, np.uint8)
    cap = L.wsb_save_compress_bound(src.size)
    out = np.empty(cap, np.uint8)
    n = L.wsb_save_compress(src.ctypes.data, src.size, out.ctypes.data, cap, level, threads)
    if n < 0:
        raise RuntimeError(f"wsb_save_compress failed ({n})")
    return out[:n].tobytes()


def inflate(stream: bytes) -> bytes:
    """pako.inflate (app.js:1267)."""
    L = native_codec()
    if L is None:
        return zlib.decompress(stream)
    src = np.frombuffer(stream, np.uint8)
    n = L.wsb_save_inflated_size(src.ctypes.data, src.size)
    if n < 0:
        raise zlib.error(f"corrupt zlib stream ({n})")
    out = np.empty(n, np.uint8)
    m = L.wsb_save_decompress(src.ctypes.data, src.size, out.ctypes.data, n)
    if m != n:
        raise zlib.error(f"corrupt zlib stream ({m})")
    return out.tobytes()

SAVE_FILE_VERSION_ID = 263574036  # app.js:345
LEGACY_VERSION_ID = 1939327491  # app.js:1265
NUM_DROPLETS_DIVIDER = 25  # app.js:452


def num_droplets(width: int, height: int) -> int:
    """app.js:1282 — JS float division; every shipped size divides evenly, otherwise the typed
    array slicing in loadData effectively floors."""
    return (width * height) // NUM_DROPLETS_DIVIDER


@dataclass
class SaveFile:
    width: int
    height: int
    base: np.ndarray  # float32 [H][W][4]
    water: np.ndarray  # float32 [H][W][4]
    wall: np.ndarray  # int8    [H][W][4]
    droplets: np.ndarray  # float32 [N][5]
    stations: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int16))
    settings_json: str | None = None  # raw guiControls JSON (None for legacy saves)
    version: int = SAVE_FILE_VERSION_ID


class IncompatibleFile(ValueError):
    """alert('Incompatible file!') at app.js:1349."""


def loads(blob: bytes) -> SaveFile:
    if len(blob) < 4:
        raise IncompatibleFile("file too short for a version id")
    (version,) = struct.unpack_from("<I", blob, 0)
    if version not in (SAVE_FILE_VERSION_ID, LEGACY_VERSION_ID):
        raise IncompatibleFile(f"unknown save file version id {version}")
    try:
        data = inflate(blob[4:])
    except zlib.error as e:
        raise IncompatibleFile(f"payload is not a zlib stream: {e}")
    L = native_codec()
    if L is not None:  # the C++ reader (csrc/wsb_save.cpp: wsb_save_parse); the Python below is its fallback and cross-check
        return _loads_native(L, data, version)
    w, h = struct.unpack_from("<HH", data, 0)
    n = w * h
    nd = num_droplets(w, h)
    off = 4
    need = off + n * 16 * 2 + n * 4 + nd * 20
    if len(data) < need:
        raise IncompatibleFile(f"payload truncated: {len(data)} < {need} bytes for {w}x{h}")
    base = np.frombuffer(data, "<f4", n * 4, off).reshape(h, w, 4).copy()
    off += n * 16
    water = np.frombuffer(data, "<f4", n * 4, off).reshape(h, w, 4).copy()
    off += n * 16
    wall = np.frombuffer(data, "i1", n * 4, off).reshape(h, w, 4).copy()
    off += n * 4
    drops = np.frombuffer(data, "<f4", nd * 5, off).reshape(nd, 5).copy()
    off += nd * 20
    stations = np.zeros((0, 2), np.int16)
    settings = None
    if version == SAVE_FILE_VERSION_ID:
        (ns,) = struct.unpack_from("<h", data, off)
        off += 2
        stations = np.frombuffer(data, "<i2", ns * 2, off).reshape(ns, 2).copy()
        off += ns * 4
        settings = data[off:].decode("utf-8")
    return SaveFile(w, h, base, water, wall, drops, stations, settings, version)


def _loads_native(L, data: bytes, version: int) -> SaveFile:
    lay = SaveLayout()
    buf = np.frombuffer(data, np.uint8)
    rc = L.wsb_save_parse(buf.ctypes.data, buf.size, version, ctypes.byref(lay))
    if rc != 0:
        raise IncompatibleFile({-4: f"unknown save file version id {version}", -5: f"payload truncated ({len(data)} bytes)"}.get(rc, f"wsb_save_parse failed ({rc})"))
    w, h, n = lay.width, lay.height, lay.width * lay.height
    base = np.frombuffer(data, "<f4", n * 4, lay.off_base).reshape(h, w, 4).copy()
    water = np.frombuffer(data, "<f4", n * 4, lay.off_water).reshape(h, w, 4).copy()
    wall = np.frombuffer(data, "i1", n * 4, lay.off_wall).reshape(h, w, 4).copy()
    drops = np.frombuffer(data, "<f4", lay.n_droplets * 5, lay.off_droplets).reshape(lay.n_droplets, 5).copy()
    stations, settings = np.zeros((0, 2), np.int16), None
    if lay.off_settings >= 0:
        stations = np.frombuffer(data, "<i2", lay.n_stations * 2, lay.off_stations).reshape(lay.n_stations, 2).copy()
        settings = data[lay.off_settings:lay.off_settings + lay.settings_len].decode("utf-8")
    return SaveFile(w, h, base, water, wall, drops, stations, settings, version)


def load(path: str) -> SaveFile:
    with open(path, "rb") as f:
        return loads(f.read())


def payload(sf: SaveFile, native: bool | None = None) -> bytes:
    """The uncompressed byte stream prepareDownload assembles (app.js:6610-6614); through the C++ writer
    (wsb_save_serialise) when libwsbsave.so is built (native=False forces the Python restatement)."""
    h, w = sf.height, sf.width
    assert sf.base.shape == (h, w, 4) and sf.water.shape == (h, w, 4) and sf.wall.shape == (h, w, 4)
    L = native_codec() if native in (None, True) else None
    if L is not None:
        st = np.ascontiguousarray(sf.stations, "<i2").reshape(-1, 2)
        js = (sf.settings_json or "{}").encode("utf-8")
        cur = sf.version == SAVE_FILE_VERSION_ID
        size = L.wsb_save_payload_size(w, h, st.shape[0] if cur else 0, len(js) if cur else 0, sf.version)
        if size < 0:
            raise ValueError(f"wsb_save_payload_size failed ({size})")
        out = np.empty(size, np.uint8)
        arrs = [np.ascontiguousarray(sf.base, "<f4"), np.ascontiguousarray(sf.water, "<f4"), np.ascontiguousarray(sf.wall, "i1"),
                np.ascontiguousarray(sf.droplets, "<f4")]
        assert arrs[3].shape == (num_droplets(w, h), 5), "droplet buffer must hold W*H/25 records"
        m = L.wsb_save_serialise(out.ctypes.data, size, w, h, *[a.ctypes.data for a in arrs], st.ctypes.data if st.size else None, st.shape[0],
                                 js, len(js), sf.version)
        if m != size:
            raise ValueError(f"wsb_save_serialise failed ({m})")
        return out.tobytes()
    parts = [
        struct.pack("<HH", w, h),
        np.ascontiguousarray(sf.base, "<f4").tobytes(),
        np.ascontiguousarray(sf.water, "<f4").tobytes(),
        np.ascontiguousarray(sf.wall, "i1").tobytes(),
        np.ascontiguousarray(sf.droplets, "<f4").tobytes(),
    ]
    if sf.version == SAVE_FILE_VERSION_ID:
        st = np.ascontiguousarray(sf.stations, "<i2").reshape(-1, 2)
        parts.append(struct.pack("<H", st.shape[0]))  # Uint16Array.of(weatherStations.length)
        parts.append(st.tobytes())
        parts.append((sf.settings_json or "{}").encode("utf-8"))
    return b"".join(parts)


def dumps(sf: SaveFile, level: int = 6, threads: int = 0) -> bytes:
    return struct.pack("<I", sf.version) + deflate(payload(sf), level, threads)


def save(path: str, sf: SaveFile, level: int = 6, threads: int = 0) -> None:
    with open(path, "wb") as f:
        f.write(dumps(sf, level, threads))
