"""Cuts 512-column windows out of two reference saves (reference checkout: saves/*.weathersandbox,
written by the reference's own WebGL path, app.js:6575-6628) and stores them in the same
.weathersandbox container under tests/golden/ — the full files are 19.5 and 25.2 MB, the crops
3-4 MB, and /root/reference does not exist on the GPU box.

    python tests/golden/make_save_crops.py [/root/reference/saves]

The window is the 512 columns centred on the save's strongest precipitation column.  A crop is a
PERIODIC 512-column domain of its own: the reference's wall / distance fields are a fixed point of
one iteration only away from the artificial seam, so the reference-output pins
(tests/test_reference_saves.py) skip SEAM_MARGIN = 136 columns on either side (DISTANCE saturates at
127); the GPU-vs-oracle parity tests use the whole crop.  Droplets inside the window keep their
place (x re-normalised to the crop), the rest of the W*H/25 slots are inactive."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import wsb200  # noqa: E402

CROP_W = 512
SAVES = {"Hotlake Valley": "hotlake_valley_crop512", "Powerful Hail and Snow Cells": "hail_and_snow_cells_crop512"}


def crop(sf, x0):
    S = wsb200.savefile
    W, H = sf.width, sf.height
    cols = np.arange(x0, x0 + CROP_W) % W
    nd = S.num_droplets(CROP_W, H)
    drops = np.zeros((nd, 5), np.float32)
    drops[:, 0] = np.random.default_rng(1).uniform(-1, 1, nd)  # inactive droplets: mass -10, random x (app.js:4901-4913)
    drops[:, 1] = -10.0
    drops[:, 2] = -10.0
    d = sf.droplets
    xg = (d[:, 0] * 0.5 + 0.5) * W
    rel = (xg - x0) % W
    keep = (d[:, 2] >= 0) & (rel < CROP_W)
    kept = d[keep][:nd].copy()
    kept[:, 0] = (rel[keep][:nd] / CROP_W) * 2.0 - 1.0
    drops[:kept.shape[0]] = kept
    return S.SaveFile(CROP_W, H, sf.base[:, cols].copy(), sf.water[:, cols].copy(), sf.wall[:, cols].copy(), drops,
                      np.zeros((0, 2), np.int16), sf.settings_json, sf.version), int(kept.shape[0])


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/saves"
    for name, out in SAVES.items():
        sf = wsb200.savefile.load(os.path.join(src, name + ".weathersandbox"))
        air = sf.wall[..., 1] != 0
        precip = np.where(air, sf.water[..., 2], 0).sum(axis=0)
        x0 = (int(np.argmax(np.convolve(np.tile(precip, 2), np.ones(CROP_W), "valid")[:sf.width]))) % sf.width
        c, n_active = crop(sf, x0)
        path = os.path.join(ROOT, "tests", "golden", out + ".weathersandbox")
        wsb200.savefile.save(path, c, level=9)
        print(f"{name}: {sf.width}x{sf.height} -> columns [{x0}, {x0 + CROP_W}) -> {path} ({os.path.getsize(path) / 1e6:.1f} MB, {n_active} active droplets)")


if __name__ == "__main__":
    main()
