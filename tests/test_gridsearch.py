"""The grid-search driver (kiwi_b200/grid_search.py, after python/tunguska/gridsearch.py MisfitGrid): grid construction and statistics
on the CPU, the whole search on the GPU against the numpy restatement of make_global_misfits."""
import numpy as np
import pytest

import scenario as sc
from oracle_outer import cube_from_block, make_global_misfits


def test_grid_construction_and_stats():
    from kiwi_b200 import grid_search as gs
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
    from kiwi_b200.grid_search import bootstrap_weights
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


def test_merge_of_sharded_results():
    from kiwi_b200.grid_search import merge_best
    nan = np.nan
    best, val = merge_best([[[3, -1, 7, -1], [0.5, nan, 0.2, nan]], [[1, 4, 2, -1], [0.5, 0.9, 0.3, nan]]])
    assert best.tolist() == [1, 4, 7, -1]                  # equal misfits: the lower candidate number, as nanargmin over the whole grid
    assert val[:3].tolist() == [0.5, 0.9, 0.2] and np.isnan(val[3])


def _gather_worker(rank, world, port, q):
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from kiwi_b200.grid_search import _gather_rows
    from kiwi_b200.sharding import balanced_partition
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    shares = balanced_partition(np.arange(7, dtype=float), world)
    local = np.stack([shares[rank] * 10.0, shares[rank] + 0.5])
    out = _gather_rows(local, shares, 7, None)
    q.put((rank, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gather_of_sharded_rows_two_ranks_gloo():
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        assert out == [[0.0, 10.0, 20.0, 30.0, 40.0, 50.0, 60.0], [0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5]], (rank, out)


def _sharded_worker(rank, world, port, q):
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    import scenario as sc_
    from kiwi_b200 import Engine, MisfitGrid
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    comps = ["ned", "ar", "d", "neu"]
    lat, lon, dep = sc_.small_receivers(4)
    g = Engine(rank)
    sc_.setup(g, sc_.small_db(), lat, lon, dep, comps)
    truth = sc_.BILAT_SMALL.copy()
    g.set_source_params("bilateral", truth)
    g.set_synthetic_reference(scale=1.05)
    ranges = [("strike", truth[5] - 30., truth[5] + 30., 10.), ("length-a", 1000., 5000., 1000.)]
    grid = MisfitGrid("bilateral", truth, param_ranges=ranges)
    grid.compute(g, costs=grid.sources[:, 9] + grid.sources[:, 10])
    grid.postprocess(bootstrap_iterations=32, seed=7)
    res = (grid.best_source.tolist(), grid.misfits_by_s.tolist(), grid.bootstrap_sources.tolist(), grid.status.tolist())
    if rank == 0:      # the same search on one GPU
        dist.barrier()
        single = MisfitGrid("bilateral", truth, param_ranges=ranges)
        g1 = Engine(0)
        sc_.setup(g1, sc_.small_db(), lat, lon, dep, comps)
        g1.set_source_params("bilateral", truth)
        g1.set_synthetic_reference(scale=1.05)
        block, status = g1.eval_sources("bilateral", single.sources)
        q.put((rank, res, block.tolist()))
    else:
        dist.barrier()
        q.put((rank, res, None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_sharded_grid_search_two_gpus():
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from kiwi_b200.grid_search import bootstrap_weights
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a, block), (r1, b, _) = res
    assert a == b                                          # every rank ends with the same result
    best, misfits, boots, status = a
    block = np.array(block, dtype=np.float64)
    m, n = cube_from_block(block, [True] * 4, [3, 2, 1, 3])
    want, _ = make_global_misfits(m, n)
    assert np.allclose(misfits, want, rtol=1e-12) and not any(status)
    bw = bootstrap_weights([True] * 4, None, 32, np.random.default_rng(7))
    from kiwi_b200.grid_search import source_grid, mimainc_to_gvals
    truth = sc.BILAT_SMALL
    grid_sources = source_grid("bilateral", truth, [("strike", mimainc_to_gvals(truth[5] - 30., truth[5] + 30., 10.)), ("length-a", mimainc_to_gvals(1000., 5000., 1000.))])
    assert np.array_equal(np.array(best, np.float32), grid_sources[np.nanargmin(want)])
    for k in range(32):
        wb, _ = make_global_misfits(m, n, bweights=bw[k][None, :])
        assert np.array_equal(np.array(boots[k], np.float32), grid_sources[np.nanargmin(wb)])
