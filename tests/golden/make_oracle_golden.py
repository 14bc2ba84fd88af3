"""Generates tests/golden/oracle_100x100.npz: the oracle's state after 1, 10 and 100 iterations
of the full loop on `100x100_test.weathersandbox` (reference saves/100 X 100 Test.weathersandbox),
parameters resolved like the reference's loader, sun pinned at the save's angle.

PARITY UNPINNED: the reference ships no expected outputs and cannot be executed here; these
vectors pin the ORACLE (against accidental change, and as the GPU tests' fixture), not the oracle
against the reference.

    python tests/golden/make_oracle_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import wsb200  # noqa: E402
from oracle import oracle as O  # noqa: E402

P = wsb200.params


def run():
    sf = wsb200.savefile.load(os.path.join(HERE, "100x100_test.weathersandbox"))
    g = P.resolve_settings(sf.settings_json)
    ora = O.OracleSim(sf.width, sf.height, sf.droplets.shape[0])
    ora.upload(sf.base, sf.water, sf.wall, sf.droplets)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(sf.height, g))
    out = {}
    done = 0
    for n in (1, 10, 100):
        ora.step(n - done)
        done = n
        out[f"base_{n}"] = ora.field(O.FIELD_BASE, 0)
        out[f"water_{n}"] = ora.field(O.FIELD_WATER, 1)
        out[f"wall_{n}"] = ora.field(O.FIELD_WALL, 0)
        out[f"light_{n}"] = ora.light_latest()
        out[f"drops_{n}"] = ora.droplets()
    ora.step(900)
    h = hashlib.sha256()
    for a in (ora.field(O.FIELD_BASE, 0), ora.field(O.FIELD_WATER, 1), ora.field(O.FIELD_WALL, 0), ora.light_latest()):
        h.update(a.tobytes())
    out["sha256_1000"] = np.frombuffer(h.digest(), np.uint8)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_100x100.npz"), **run())
    print("written", os.path.join(HERE, "oracle_100x100.npz"))
