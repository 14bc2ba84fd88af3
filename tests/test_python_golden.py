"""tests/golden/python_100x100.npz holds the state of the reference's own 100 x 100 save after 1, 10
and 101 iterations of the full loop (particles included) as computed by the independent Python
restatements of the shaders (tests/golden/make_python_golden.py) — not by the oracle and not by
the CUDA path.  The oracle must hit these vectors bit for bit; so must the oracle-generated
fixture the GPU tests use (where the two overlap), which pins the GPU golden test to the same
independently generated vectors; and the fused kernels on the host emulator must stay within the
north-star tolerance (their additive sprites sum in a different order)."""
import os

import numpy as np
import pytest

import wsb200
from oracle import oracle as O

P = wsb200.params
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "python_100x100.npz"))


@pytest.fixture(scope="module")
def save():
    sf = wsb200.savefile.load(os.path.join(GOLDEN, "100x100_test.weathersandbox"))
    return sf, P.resolve_settings(sf.settings_json)


def test_oracle_reproduces_the_python_generated_vectors(gold, save):
    sf, g = save
    ora = O.OracleSim(sf.width, sf.height, sf.droplets.shape[0])
    ora.upload(sf.base, sf.water, sf.wall, sf.droplets)
    ora.set_params(P.derive_params(g))
    ora.set_frame_inputs(P.frame_inputs(g))
    ora.set_profiles(P.initial_T_profile(sf.height, g))
    done = 0
    for n in (1, 10, 101):
        ora.step(n - done)
        done = n
        for name, got in (("base", ora.field(O.FIELD_BASE, 0)), ("water", ora.field(O.FIELD_WATER, 1)), ("wall", ora.field(O.FIELD_WALL, 0)),
                          ("light", ora.light_latest()), ("drops", ora.droplets()), ("feedback", ora.field(O.FIELD_FEEDBACK)),
                          ("deposition", ora.field(O.FIELD_DEPOSITION))):
            want = gold[f"{name}_{n}"]
            assert np.array_equal(got, want), f"{name} after {n} iterations: {(got != want).sum()} values differ from the Python-generated vectors"


def test_the_two_fixtures_agree_where_they_overlap(gold):
    """oracle_100x100.npz (oracle-generated, the GPU tests' fixture) == python_100x100.npz at 1 and 10 iterations."""
    og = np.load(os.path.join(GOLDEN, "oracle_100x100.npz"))
    for n in (1, 10):
        for name in ("base", "water", "wall", "light", "drops"):
            assert np.array_equal(og[f"{name}_{n}"], gold[f"{name}_{n}"]), f"{name}_{n}"


def test_fused_kernels_on_the_emulator_track_the_python_generated_vectors(gold, save, emu):
    from test_host_cells import EmuFused, _ptr

    sf, g = save
    em = EmuFused(emu, g, sf.base, sf.water, sf.wall)
    drops = np.ascontiguousarray(sf.droplets)
    emu.ef_upload_drops(em.h, _ptr(drops), drops.shape[0])
    done = 0
    for n in (1, 10):
        emu.ef_step(em.h, n - done)
        done = n
        assert np.array_equal(em.read(2, 0), gold[f"wall_{n}"])
        for name, f, v in (("base", 0, 0), ("water", 1, 1), ("light", 3, 2)):
            got, want = em.read(f, v).astype(np.float64), gold[f"{name}_{n}"].astype(np.float64)
            err = np.max(np.abs(got - want) / (np.abs(want) + 1e-3))
            assert err < 1e-5, f"{name} after {n} iterations: relative error {err:.3g}"
        d = np.empty_like(drops)
        emu.ef_read_drops(em.h, _ptr(d))
        assert np.array_equal(d[:, 2] < 0, gold[f"drops_{n}"][:, 2] < 0)
    em.close()
