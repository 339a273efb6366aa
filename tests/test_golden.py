"""Committed fixtures (tests/golden/small_scenario.npz, generator tests/golden/make_golden.py): the oracle
must keep reproducing them bit for bit (CPU), the CUDA path must match them within the parity bars (GPU)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "small_scenario.npz"))


def test_oracle_reproduces_golden_fixtures_bit_exact():
    now = mg.build()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        a, b = np.asarray(now[k]), GOLD[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b), k


@pytest.mark.gpu
def test_cuda_path_against_golden_fixtures():
    from kiwi_b200 import Engine
    now = mg.build(lambda: Engine(0))
    for k in GOLD.files:
        a, b = np.asarray(now[k]), GOLD[k]
        if k.startswith(("table_", "grid_")) or k.endswith("_first"):
            assert np.array_equal(a, b), k                                   # discretisation and spans: exact
        elif k.startswith("seis_"):
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), k
        else:
            floor = 0.25 if "ampspec" in k else 0.1
            tol = 1e-5 * np.maximum(np.abs(b), floor * np.abs(b[..., 1:2]))
            assert np.all(np.abs(a - b) <= tol), (k, np.abs((a - b) / tol).max())


EXTRAS = np.load(os.path.join(HERE, "golden", "round1_extras.npz"))


def test_oracle_reproduces_the_round1_extras_bit_exact():
    now = mg.build_extras()
    assert sorted(now) == sorted(EXTRAS.files)
    for k in EXTRAS.files:
        a, b = np.asarray(now[k]), EXTRAS[k]
        assert a.shape == b.shape and a.dtype == b.dtype, (k, a.dtype, b.dtype)
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b), k


@pytest.mark.gpu
def test_cuda_path_against_the_round1_extras():
    from kiwi_b200 import Engine
    now = mg.build_extras(lambda: Engine(0))
    for k in EXTRAS.files:
        a, b = np.asarray(now[k]), EXTRAS[k]
        assert a.shape == b.shape, k
        if k.endswith("_first") or k in ("autoshift", "distances", "azimuths", "principal_axes", "spectrum_df") or k.startswith("interp_"):
            assert np.array_equal(a, b), k                     # integers, host arithmetic, the bit-exact interpolation
        elif k.startswith("misfits_"):
            tol = 1e-5 * np.maximum(np.abs(b), 0.25 * np.abs(b[..., 1:2]))      # filtered norm: through two fp32 FFTs
            assert np.all(np.abs(a - b) <= tol), (k, np.abs((a - b) / tol).max())
        else:                                                  # cross-correlations, probes, spectrum
            assert np.abs(a - b).max() <= 4e-5 * np.abs(b).max(), (k, np.abs(a - b).max() / np.abs(b).max())
