"""Inputs the reference accepts without a bound and round 1 of this engine refused: more than 32 time centroids per
sub-fault (source_bilat.f90:274-315), rise-time folds of more than 1024 shifts (sparse_trace.f90:379-402,
receiver.f90:868-885), padded probe spans beyond 16384 samples (comparator.f90:1092-1118) and synthetic windows of several
thousand samples (long traces of a teleseismic database).  Each against the oracle, through the C ABI."""
import numpy as np
import pytest

import scenario as sc
from oracle_lib import OracleEngine
from test_parity_gpu import COMPS6, EIK, RTOL, assert_seis_close, engines, misfit_tol

pytestmark = pytest.mark.gpu


def test_48_time_centroids_per_sub_fault():
    """bilateral source with a long rise time: nt = floor((rise + len/nx/v) / dt_eff) + 1 = 48 > 32"""
    p = sc.BILAT_SMALL.copy()
    p[13] = 9.7                       # rise time [s]; effective_dt = 0.2 (scenario.setup)
    g, o = engines(sc.small_db(), COMPS6)
    tg, gg, ng = g.discretize_source("bilateral", p)
    to, go, no = o.discretize_source("bilateral", p)
    assert list(gg) == list(go) and gg[2] >= 48 and ng == no
    assert np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    o.eval_sources("bilateral", p)
    g.set_source_params("bilateral", p)
    # ~1e4 centroids: the fp32 restatement's own accumulation noise is at the bar here (1.4e-5 of the peak against the GPU), so the
    # restatement with the strips carried in double arbitrates, as in tests/test_fullsize_parity_gpu.py
    lat, lon, dep = sc.small_receivers(6)
    ow = OracleEngine(wide=True)
    sc.setup(ow, sc.small_db(), lat, lon, dep, COMPS6)
    ow.eval_sources("bilateral", p)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            assert_seis_close(g.get_seismogram(ir, ic), ow.get_seismogram(ir, ic), "rcv %d comp %d" % (ir, ic))
            (fg, dg), (fo, do), (fw, dw) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic), ow.get_seismogram(ir, ic)
            noise = np.abs(do - dw).max()           # the fp32 restatement's own distance from the exactly accumulated sum
            assert (fg, dg.size) == (fo, do.size) and np.abs(dg - do).max() <= RTOL * np.abs(do).max() + 1.5 * noise
    sc.set_refs_from(o, [g, o, ow], [len(c) for c in COMPS6])
    q = np.tile(p, (3, 1)); q[1, 13] = 14.0; q[2, 5] += 10      # 71 time centroids; another strike
    mg, sg = g.eval_sources("bilateral", q)
    mo, so = o.eval_sources("bilateral", q)
    mw, sw = ow.eval_sources("bilateral", q)
    assert np.array_equal(sg, so) and not sg.any()
    assert np.all(np.abs(mg - mw) <= misfit_tol(mw))
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo) + 1.5 * np.abs(mo - mw))


def test_rise_time_fold_of_1500_shifts():
    """eikonal source with a rise time of 150 s at dt = 0.1: 1501 shifted copies in strip_fold"""
    p = EIK.copy()
    p[14] = 150.0
    g, o = engines(sc.small_db(), COMPS6)
    o.eval_sources("eikonal", p)
    g.set_source_params("eikonal", p)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic, 1), o.get_seismogram(ir, ic, 1)
            # (the end of the folded strip is decided by fp32 noise in the reference: test_eikonal_seismograms_with_rise_time_fold)
            assert fg == fo and abs(dg.size - do.size) <= 8 and dg.size > 1500, (ir, ic, fg, fo, dg.size, do.size)
            n = max(dg.size, do.size)
            eg = np.concatenate([dg, np.full(n - dg.size, dg[-1], np.float32)])
            eo = np.concatenate([do, np.full(n - do.size, do[-1], np.float32)])
            assert np.abs(eg - eo).max() <= RTOL * np.abs(eo).max(), (ir, ic, np.abs(eg - eo).max() / np.abs(eo).max())


@pytest.mark.parametrize("norm", ["ampspec_l2norm", "l2norm"])
def test_probe_span_of_32768_samples(norm):
    """references of 9000 samples: the padded span of the probes is 32768 (amplitude spectra / band-pass filtered norms through
    the global-memory transform buffer)"""
    g, o = engines(sc.small_db(), COMPS6)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    rng = np.random.default_rng(3)
    nc = [len(c) for c in COMPS6]
    for ir in range(1, 7):
        for ic in range(1, nc[ir - 1] + 1):
            first, data = o.get_seismogram(ir, ic, 1)
            ref = np.zeros(9000, np.float32)
            ref[:data.size] = data * np.float32(1.07)
            ref += (np.abs(data).max() * 0.05 * rng.standard_normal(9000)).astype(np.float32)
            for e in (g, o):
                e.set_ref_seismogram(ir, ic, (first - 1) * 0.1, ref)
    for e in (g, o):
        e.set_misfit_method(norm)
        if norm == "l2norm":
            e.set_misfit_filter([0.05, 0.1, 1.0, 2.0], [0, 1, 1, 0])
    p = np.tile(sc.BILAT_SMALL, (2, 1)); p[1, 5] += 20
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    assert np.array_equal(sg, so) and not sg.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo, 0.25)), np.abs((mg - mo) / misfit_tol(mo, 0.25)).max()


def long_trace_db():
    """4000-sample traces on a small grid (a stand-in for a teleseismic database): smooth random wavelets, deterministic"""
    from kiwi_b200 import Gfdb
    rng = np.random.default_rng(11)
    nx, nz, n = 24, 6, 4000
    db = Gfdb.create(nx, nz, 10, 0.5, 1000.0, 1000.0, 20000.0, 0.0)
    t = np.arange(n, dtype=np.float64)
    for ix in range(1, nx + 1):
        for iz in range(1, nz + 1):
            onset = 40.0 + 3.0 * ix + 1.5 * iz
            for ig in range(1, 11):
                f = 0.002 + 0.0005 * ig
                tr = np.sin(2 * np.pi * f * (t - onset) + ig) * np.exp(-((t - onset - 1500.0) / 1200.0) ** 2) * (t > onset)
                tr += 0.1 * rng.standard_normal(n) * (t > onset)
                db.save_array(ix, iz, ig, 100, tr.astype(np.float32))
    return db


def test_synthetic_window_of_4000_samples():
    """windows far beyond the 500 samples the shared-memory accumulators of k_synth hold at full occupancy"""
    from kiwi_b200 import Engine
    db = long_trace_db()
    lat, lon, dep = sc.small_receivers(4, dmin=26e3, dmax=36e3)
    comps = ["ned", "ar", "d", "neu"]
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, effective_dt=1.0)
    p = np.array([1.0, 500, -800, 2500, 1.5e18, 75, 70, 150, 20, 3000, 2000, 2000, 3000, 2.0], dtype=np.float32)
    o.eval_sources("bilateral", p)
    g.set_source_params("bilateral", p)
    for ir in range(1, 5):
        for ic in range(1, len(comps[ir - 1]) + 1):
            fg, dg = g.get_seismogram(ir, ic)
            assert dg.size > 3900
            assert_seis_close((fg, dg), o.get_seismogram(ir, ic), "rcv %d comp %d" % (ir, ic))
    sc.set_refs_from(o, [g, o], [len(c) for c in comps], dt=0.5)
    q = np.tile(p, (3, 1)); q[1, 6] -= 15; q[2, 3] += 400
    mg, sg = g.eval_sources("bilateral", q)
    mo, so = o.eval_sources("bilateral", q)
    assert np.array_equal(sg, so) and not sg.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo))


def test_moment_tensor_grid_beyond_the_fused_kernels_limits():
    """a point moment-tensor grid whose rise time gives more distinct quad shifts than the fused kernel (k_mt_fused) keeps: the engine
    falls back to the unfused tensor-core path (k_synth + k_mt_contract) and says nothing; results against the oracle as for any other
    grid (second pass: a short rise time, fused kernel, no synthesis stage of its own)"""
    from test_parity_gpu import _mt_grid
    ncomps = [len(c) for c in COMPS6]
    for risetime in (9.0, 0.7):
        g, o = engines(sc.small_db(), COMPS6)
        o.eval_sources("moment_tensor", sc.MT_SMALL)
        sc.set_refs_from(o, [g, o], ncomps)
        p = _mt_grid()
        p[:, 10] = risetime                   # 9 s at effective_dt 0.2: 46 time centroids, 12 distinct quad shifts
        mg, sg = g.eval_sources("moment_tensor", p)
        t = g.last_timing()["launches"]
        assert t[3] >= 1                      # a tensor-core path ran
        assert (t[2] >= 1) == (risetime > 5)  # ... with a synthesis stage of its own only where the fused kernel does not fit
        mo, so = o.eval_sources("moment_tensor", p)
        assert not sg.any() and not so.any()
        assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()


def test_moment_tensor_grid_on_4000_sample_windows():
    """the same on the long-trace database: the strips of a (location, receiver) pair do not fit the fused kernel's shared memory"""
    from kiwi_b200 import Engine
    from kiwi_b200 import synthetic
    db = long_trace_db()
    lat, lon, dep = sc.small_receivers(4, dmin=26e3, dmax=36e3)
    comps = ["ned", "ar", "d", "neu"]
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, effective_dt=1.0)
    base = np.array([1.0, 500, -800, 2500, 1e18, -0.4e18, -0.6e18, 0.3e18, 0.2e18, -0.5e18, 2.0], dtype=np.float32)
    o.eval_sources("moment_tensor", base)
    sc.set_refs_from(o, [g, o], [len(c) for c in comps], dt=0.5)
    mts = synthetic.fibonacci_moment_tensors(40) * 1e18
    p = np.tile(base, (2 * 40, 1))
    p[:, 4:10] = np.concatenate([mts, mts])
    p[40:, 1] += 700; p[40:, 3] += 500
    mg, sg = g.eval_sources("moment_tensor", p)
    t = g.last_timing()["launches"]
    assert t[3] >= 1 and t[2] >= 1           # tensor-core path, unfused
    mo, so = o.eval_sources("moment_tensor", p)
    assert not sg.any() and not so.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()
