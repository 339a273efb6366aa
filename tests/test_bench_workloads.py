"""The synthetic workloads of bench.py (CPU only): every candidate of a step is a different source, so that no evaluation of the timed
region is answered from another one's synthesis (the engine shares syntheses between candidates that differ only in the moment)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_candidates_of_a_step_are_distinct_sources(bench, name):
    w = bench.WORKLOADS[name]
    stype, cands, base = bench.candidates(w, max(w["batch"], 32))
    step = np.ascontiguousarray(cands[: w["batch"]])
    assert step.shape[0] == w["batch"]
    geometry = np.delete(step, 4, axis=1)                    # everything but the moment
    assert np.unique(geometry, axis=0).shape[0] == w["batch"], "candidates that differ only in the moment would share one synthesis"
    assert np.isfinite(step).all()


def test_moment_tensor_grid_has_many_tensors_per_location(bench):
    w = bench.WORKLOADS["c2"]
    stype, cands, base = bench.candidates(w, w["batch"])
    assert stype == "moment_tensor" and cands.shape == (w["batch"], 11)
    loc = np.concatenate([cands[:, :4], cands[:, 10:11]], axis=1)
    nloc = np.unique(loc, axis=0).shape[0]
    assert nloc * 8 <= w["batch"]                            # the grid path of the engine applies (engine.cpp eval_mt_grid)
    assert np.unique(cands, axis=0).shape[0] == w["batch"]   # and no candidate is listed twice
