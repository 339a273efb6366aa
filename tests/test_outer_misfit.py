"""Outer misfit / bootstrap / best source: the device reduction against the numpy
restatement of make_global_misfits (tests/oracle_outer.py)."""
import numpy as np
import pytest

import scenario as sc
from oracle_outer import cube_from_block, make_global_misfits

COMPS = ["ned", "ar", "d", "neu", "cl", "wsd"]


def test_restatement_matches_the_engine_formula_for_plain_l2():
    """minimizer_engine.f90:939-942 is make_global_misfits(l2norm) with unit weights"""
    rng = np.random.default_rng(1)
    m = rng.uniform(0.1, 1, (3, 4, 3)); n = rng.uniform(0.5, 2, (3, 4, 3))
    g, _ = make_global_misfits(m, n)
    want = np.sqrt((m ** 2).sum((1, 2))) / np.sqrt((n ** 2).sum((1, 2)))
    assert np.allclose(g, want, rtol=1e-14)
    g1, _ = make_global_misfits(m, n, outer_norm="l1norm")
    assert np.allclose(g1, m.sum((1, 2)) / n.sum((1, 2)), rtol=1e-14)
    ga, _ = make_global_misfits(m, n, anarchy=True)       # every receiver normalised by its own norm
    per = np.sqrt((m ** 2).sum(2)) / np.sqrt((n ** 2).sum(2))
    assert np.allclose(ga, np.sqrt((per ** 2).sum(1) / 4.0), rtol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("outer_norm", ["l2norm", "l1norm"])
@pytest.mark.parametrize("anarchy", [False, True])
def test_outer_misfits_on_device(outer_norm, anarchy, monkeypatch):
    from kiwi_b200 import Engine
    from oracle_lib import OracleEngine
    lat, lon, dep = sc.small_receivers(6)
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, sc.small_db(), lat, lon, dep, COMPS)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g], [len(c) for c in COMPS])
    g.switch_receiver(3, False)
    p = np.tile(sc.BILAT_SMALL, (9, 1))
    p[:, 5] += np.linspace(-40, 40, 9); p[4, 5] = sc.BILAT_SMALL[5] + 3
    p[7, 12] = -1.0                                   # negative rupture velocity: this candidate fails
    block, status = g.eval_sources("bilateral", p)    # also leaves the cube on the device
    assert status[7] == 1 and status.sum() == 1
    enabled = [True, True, False, True, True, True]
    ncomps = [len(c) for c in COMPS]
    m, n = cube_from_block(block.astype(np.float64), enabled, ncomps)
    rng = np.random.default_rng(5)
    weights = rng.uniform(0.5, 2.0, 6)
    nboot = 16
    idx = np.array([0, 1, 3, 4, 5])
    bw = np.zeros((nboot, 6))
    for b in range(nboot):                            # seismosizer.py:869-873
        cnt = np.bincount(idx[rng.integers(0, 5, 5)], minlength=6)
        bw[b] = cnt
    out, best, bestv = g.outer_misfits(receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bweights=bw)
    want0, _ = make_global_misfits(m, n, receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy)
    assert np.isnan(out[0, 7]) and np.isnan(want0[7])
    ok = ~np.isnan(want0)
    assert np.allclose(out[0][ok], want0[ok], rtol=1e-12)
    assert best[0] == np.nanargmin(want0) == 4
    for b in range(nboot):
        wb, _ = make_global_misfits(m, n, receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bweights=bw[b][None, :])
        okb = ~np.isnan(wb)
        assert np.array_equal(np.isnan(out[b + 1]), np.isnan(wb))
        assert np.allclose(out[b + 1][okb], wb[okb], rtol=1e-12)
        assert best[b + 1] == np.nanargmin(wb) and abs(bestv[b + 1] - np.nanmin(wb)) <= 1e-12 * np.nanmin(wb)
    # the cube never has to leave the GPU: evaluate without download, reduce, same numbers
    st2 = g.eval_sources_on_device("bilateral", p)
    out2, best2, _ = g.outer_misfits(receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bweights=bw, want_matrix=True)
    assert np.array_equal(st2, status) and np.array_equal(best2, best) and np.allclose(out2[:, ok], out[:, ok], rtol=0, atol=0)
    # only the minima wanted: the bootstrap rows are reduced in passes (here three rows at a time), the matrix never exists as a whole
    monkeypatch.setenv("KIWI_OUTER_PASS_BYTES", str(3 * 9 * 8))
    out3, best3, bestv3 = g.outer_misfits(receiver_weights=weights, outer_norm=outer_norm, anarchy=anarchy, bweights=bw, want_matrix=False)
    assert out3 is None and np.array_equal(best3, best) and np.array_equal(bestv3, bestv)


@pytest.mark.gpu
def test_outer_misfits_refuses_a_stale_device_block():
    """the misfit block kiwi_eval_sources leaves on the device is gone after anything that rewrites it (a single evaluation
    behind get_misfits, minimize_lm) or changes the receivers; kiwi_outer_misfits(NULL) must say so instead of reducing garbage"""
    from kiwi_b200 import Engine, KiwiError
    from oracle_lib import OracleEngine
    lat, lon, dep = sc.small_receivers(6)
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, sc.small_db(), lat, lon, dep, COMPS)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g], [len(c) for c in COMPS])
    p = np.tile(sc.BILAT_SMALL, (4, 1)); p[:, 5] += np.arange(4) * 5.0
    g.eval_sources_on_device("bilateral", p)
    out, best, _ = g.outer_misfits(4)
    g.set_source_params("bilateral", p[2]); g.get_misfits()            # a single evaluation reuses the buffer
    with pytest.raises(KiwiError, match="no misfit block"):
        g.outer_misfits(4)
    g.eval_sources_on_device("bilateral", p)
    g.switch_receiver(2, False)                                        # the layout of the block no longer matches the receivers
    with pytest.raises(KiwiError, match="no misfit block"):
        g.outer_misfits(4)
    g.eval_sources_on_device("bilateral", p)
    out2, best2, _ = g.outer_misfits(4)
    assert best2[0] == best[0] or np.isfinite(out2).all()
