"""Config C1 (BASELINE.json configs[0]): the reference's own benchmark input benchmark/mini.inp.
Block 1 (lines 1-17): Izmit bilateral source, the 11 receivers of benchmark/izmit-receivers.table (587-1442 km),
effective_dt 0.5, bilinear -- on an analytical full-space database large enough for those distances, because the
Gemini database it names is not shipped.  Block 2 (lines 19-25): `benchdb` = the kiwibench recipe
(benchmark/kiwibench.py:45-91), effective_dt 0.1; its receivers.table is not shipped, receivers are generated."""
import functools

import numpy as np
import pytest

import scenario as sc
from kiwi_b200 import Gfdb, synthetic
from oracle_lib import OracleEngine

# benchmark/izmit-receivers.table (lat lon components)
IZMIT_RECEIVERS = [(42.350, 13.400), (49.780, 17.540), (45.490, 25.950), (47.920, 19.890), (35.870, 14.520), (34.960, 33.330),
                   (35.280, 24.890), (35.180, 25.500), (49.630, 22.710), (36.370, 25.460), (42.620, 23.240)]
SRC_91 = np.array([0, 0, 0, 10000, 2e20, 91, 87, 164, 0, 20000, 10000, 9000, 3500, 2], np.float32)      # mini.inp:6
SRC_92 = np.array([0, 0, 0, 10000, 2e20, 92, 87, 164, 0, 20000, 10000, 9000, 3500, 2], np.float32)      # mini.inp:8
SRC_BENCHDB = np.array([0, 0, 0, 5000, 1.0, 91, 87, 164, 0, 900, 700, 1000, 2500, 0.2], np.float32)     # mini.inp:24


@functools.lru_cache(maxsize=None)
def izmit_db():
    # 1000 m ... 1500 km x 0 ... 19 km, dt 0.5 s ("for C1's Izmit receivers use dx=1000 m, nx=1500, dt=0.5 s")
    return Gfdb.create(1500, 20, 10, 0.5, 1000.0, 1000.0, 1000.0, 0.0).build_ahfull(2700.0, 6000.0, 3464.0)


def setup_block1(e):
    e.set_database(izmit_db())
    e.set_effective_dt(0.5)
    e.set_local_interpolation("bilinear")
    lat, lon = zip(*IZMIT_RECEIVERS)
    e.set_receivers(lat, lon, np.zeros(11, np.float32), ["ned"] * 11)
    e.set_source_location(40.75, 29.86, 0.0)


def test_c1_grid_sizes_on_the_oracle():
    o = OracleEngine()
    setup_block1(o)
    t, grid, n = o.discretize_source("bilateral", SRC_91)
    assert list(grid) == [35, 6, 5] and n == 1050                      # C1 shorthand
    o2 = OracleEngine()
    o2.set_database(synthetic.bench_s_db(40, 40))
    o2.set_effective_dt(0.1)
    o2.set_source_location(30.0, 70.0, 0.0)
    t, grid, n = o2.discretize_source("bilateral", SRC_BENCHDB)
    assert list(grid) == [13, 5, 3] and n == 195


@pytest.mark.gpu
def test_c1_block1_izmit_receivers():
    from kiwi_b200 import Engine
    g, o, ow = Engine(0), OracleEngine(), OracleEngine(wide=True)
    for e in (g, o, ow):
        setup_block1(e)
    tg, gg, ng = g.discretize_source("bilateral", SRC_91)
    to, go, no = o.discretize_source("bilateral", SRC_91)
    assert ng == no == 1050 and np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    # output_seismograms ... synthetics plain: every trace of every receiver
    # 1050 centroids x 10 components summed sequentially in fp32: the reference path's own accumulation noise is
    # close to 1e-5 here, so the traces are also held against the restatement with strips carried in double
    o.set_source_params("bilateral", SRC_91)
    ow.set_source_params("bilateral", SRC_91)
    g.set_source_params("bilateral", SRC_91)
    for ir in range(1, 12):
        for ic in range(1, 4):
            (fg, dg), (fo, do), (fw, dw) = g.get_seismogram(ir, ic, 1), o.get_seismogram(ir, ic, 1), ow.get_seismogram(ir, ic, 1)
            assert (fg, dg.size) == (fo, do.size)
            peak = np.abs(dw).max()
            noise = np.abs(do - dw).max() / peak
            assert np.abs(dg - dw).max() <= 1e-5 * peak, (ir, ic)
            assert np.abs(dg - do).max() <= max(1e-5, 2 * noise) * peak, (ir, ic, noise)
    # L2 misfit of the alternating sources of mini.inp against each other (references = the 91-degree source)
    sc.set_refs_from(ow, [g, o, ow], [3] * 11, scale=1.0, dt=0.5)
    for e in (g, o, ow):
        e.set_misfit_method("l2norm")
    pair = np.stack([SRC_92, SRC_91])
    mg, sg = g.eval_sources("bilateral", pair)
    mo, so = o.eval_sources("bilateral", pair)
    mw, sw = ow.eval_sources("bilateral", pair)
    assert not sg.any() and not so.any() and not sw.any()
    dev = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 0.1 * np.abs(b[..., 1:2]))))
    d32, dw, d32w = dev(mg[0], mo[0]), dev(mg[0], mw[0]), dev(mo[0], mw[0])
    assert dw <= 1e-5 and d32 <= max(1e-5, 2.0 * d32w), (d32, dw, d32w)
    from kiwi_b200 import global_misfits
    assert global_misfits(mg[1:2])[0] <= 1e-5          # the reference source itself
    assert abs(global_misfits(mg[0:1])[0] - global_misfits(mw[0:1])[0]) <= 1e-5 * global_misfits(mw[0:1])[0]


@pytest.mark.gpu
def test_c1_block2_kiwibench_database():
    from kiwi_b200 import Engine
    db = synthetic.bench_s_db()                                        # gfdb_build benchdb 1 200 200 10 0.1 50 50 50 0
    lat, lon, dep = synthetic.receivers(10, (30.0, 70.0), 3000.0, 8000.0, seed=4)
    dep = np.linspace(0, 400, 10).astype(np.float32)                   # receivers.table has_depth
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        e.set_database(db); e.set_effective_dt(0.1); e.set_local_interpolation("bilinear")
        e.set_receivers(lat, lon, dep, ["ned"] * 10); e.set_source_location(30.0, 70.0, 0.0)
    tg, gg, ng = g.discretize_source("bilateral", SRC_BENCHDB)
    to, go, no = o.discretize_source("bilateral", SRC_BENCHDB)
    assert list(gg) == [13, 5, 3] and ng == no == 195 and np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    o.set_source_params("bilateral", SRC_BENCHDB)
    for ir in range(1, 11):
        for ic in range(1, 4):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic, 0), o.get_seismogram(ir, ic, 0)
            assert (fg, dg.size) == (fo, do.size)
            assert np.abs(dg - do).max() <= 1e-5 * np.abs(do).max(), (ir, ic)
