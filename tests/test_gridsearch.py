"""The grid-search driver (kiwi_b200/gridsearch.py, after python/tunguska/gridsearch.py MisfitGrid): grid construction and statistics
on the CPU, the whole search on the GPU against the numpy restatement of make_global_misfits."""
import numpy as np
import pytest

import scenario as sc
from oracle_outer import cube_from_block, make_global_misfits


def test_grid_construction_and_stats():
    from kiwi_b200 import gridsearch as gs
    assert np.allclose(gs.mimainc_to_gvals(0., 1., 0.3), [0., 1. / 3, 2. / 3, 1.])       # the increment is adjusted (gridsearch.py:18-22)
    assert gs.mimainc_to_gvals(5., 5., 1.).tolist() == [5.]
    g = gs.source_grid("bilateral", sc.BILAT_SMALL, [("strike", [10., 20., 30.]), ("depth", [1e3, 2e3])])
    assert g.shape == (6, 14)
    assert g[:, 5].tolist() == [10., 10., 20., 20., 30., 30.] and g[:, 3].tolist() == [1e3, 2e3] * 3     # first parameter slowest (source.py:139-175)
    assert np.array_equal(g[:, [0, 1, 2, 4]], np.tile(sc.BILAT_SMALL[[0, 1, 2, 4]], (6, 1)))
    with pytest.raises(ValueError):
        gs.source_grid("bilateral", sc.BILAT_SMALL, [("no-such-parameter", [1.])])
    st = gs.MisfitGridStats("strike", 20., [10., 20., 20., 20., 30.], tested_values=np.array([10., 20., 30.]))
    assert st.median == 20. and st.percentile16 == pytest.approx(np.percentile([10, 20, 20, 20, 30], 16.) - 5.) and not st.percentile84_warn
    assert "Strike = 20" in st.str_best_and_confidence()
    rng = np.random.default_rng(0)
    bw = gs.bootstrap_weights([True, False, True, True], np.array([1., 1., 0., 2.]), 50, rng)
    assert bw.shape == (50, 4) and not bw[:, 1].any() and not bw[:, 2].any() and np.all(bw.sum(1) == 2)      # only receivers 1 and 4 are usable
    blk = np.array([[3., 4.], [4., 3.], [1., 2.]])
    assert gs.global_misfit_of_one_source(blk, ["ne", "d"], [True, True]) == pytest.approx(np.sqrt(26.) / np.sqrt(29.))
    assert gs.global_misfit_of_one_source(blk, ["ne", "d"], [True, True], outer_norm="l1norm") == pytest.approx(8. / 9.)


@pytest.mark.gpu
@pytest.mark.parametrize("outer_norm,anarchy", [("l2norm", False), ("l1norm", True)])
def test_grid_search_against_the_restatement(outer_norm, anarchy):
    from kiwi_b200 import Engine, MisfitGrid
    from kiwi_b200.gridsearch import bootstrap_weights
    from oracle_lib import OracleEngine
    comps = ["ned", "ar", "d", "neu", "cl", "wsd"]
    lat, lon, dep = sc.small_receivers(6)
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, sc.small_db(), lat, lon, dep, comps)
    truth = sc.BILAT_SMALL.copy()
    o.eval_sources("bilateral", truth)
    sc.set_refs_from(o, [g], [len(c) for c in comps], scale=1.0)
    g.switch_receiver(5, False)
    grid = MisfitGrid("bilateral", truth, param_ranges=[("strike", truth[5] - 30., truth[5] + 30., 10.), ("depth", truth[3] - 1000., truth[3] + 1000., 500.)])
    assert grid.sources.shape == (35, 14)
    grid.compute(g)
    weights = np.array([1., 2., 0.5, 1., 1., 0.])                       # receiver 6 weighted out, receiver 5 disabled
    best = grid.postprocess(receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bootstrap_iterations=64, seed=3)
    assert best[5] == truth[5] and best[3] == truth[3] and grid.get_best_misfit() < 1e-5        # the true source is on the grid
    assert grid.ref_misfit < 1e-5
    # the same with the numpy restatement on the downloaded cube and the same bootstrap draws
    block, status = g.eval_sources("bilateral", grid.sources)
    enabled = [True, True, True, True, False, True]
    m, n = cube_from_block(block.astype(np.float64), enabled, [len(c) for c in comps])
    want, _ = make_global_misfits(m, n, receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy)
    assert np.allclose(grid.misfits_by_s, want, rtol=1e-12, equal_nan=True)
    bw = bootstrap_weights(enabled, weights, 64, np.random.default_rng(3))
    for b in range(64):
        wb, _ = make_global_misfits(m, n, receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bweights=bw[b][None, :])
        assert np.array_equal(grid.bootstrap_sources[b], grid.sources[np.nanargmin(wb)])
    st = grid.stats["strike"]
    assert st.best == truth[5] and st.distribution.size == 64 and st.percentile16 <= truth[5] <= st.percentile84


@pytest.mark.gpu
def test_synthetic_reference_and_receivers_snapshot():
    """set_synthetic_reference (seismosizer.py:523-527) and get_receivers_snapshot (:541-610) on the in-memory getters"""
    from kiwi_b200 import Engine
    comps = ["ned", "ar", "d"]
    lat, lon, dep = sc.small_receivers(3)
    g = Engine(0)
    sc.setup(g, sc.small_db(), lat, lon, dep, comps)
    g.set_source_params("bilateral", sc.BILAT_SMALL)
    g.switch_receiver(2, False)
    g.set_synthetic_reference()
    g.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
    assert g.get_global_misfit() < 1e-6                       # the references are the synthetics (both go through one complex FFT)
    p2 = sc.BILAT_SMALL.copy(); p2[5] += 25
    g.set_source_params("bilateral", p2)
    assert g.get_global_misfit() > 0.05
    snap = g.get_receivers_snapshot()
    assert snap[1] is None and len(snap[0]["syn_seismograms"]) == 3 and len(snap[2]["ref_spectra"]) == 1
    t0, dt, syn = snap[0]["syn_seismograms"][0]
    df, amp = snap[0]["syn_spectra"][0]
    assert dt == pytest.approx(0.1) and syn.size > 20 and amp.size > 16 and df > 0 and np.argmax(amp) * df < 3.0
    r0, _, ref = snap[0]["ref_seismograms"][0]
    assert ref.size > 20 and not np.array_equal(ref[:10], syn[:10])
