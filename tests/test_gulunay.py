"""Gulunay f-k interpolation of the database (interpolation.f90, gfdb.f90:1109-1310).  The reference has no test
of it and FFTW is absent: the oracle is checked on plane waves (what the method is exact for), the CUDA path against the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from kiwi_b200 import Gfdb

T = 256


def plane_wave_2d(s, slowness, x):
    tt = np.arange(T)[None, :]
    return np.exp(-0.5 * ((tt - 60 - slowness * x[:, None]) / 6.0) ** 2).astype(np.float32)


def test_oracle_gulunay2d_interpolates_plane_waves():
    s = 64
    for slowness in (0.5, 1.5):
        a = plane_wave_2d(s, slowness, np.arange(s, dtype=np.float64))[None]
        tapered, out = ol.gulunay(a, 2, 1, 20, 16, 0)
        true = plane_wave_2d(2 * s, slowness, np.arange(2 * s) / 2.0)
        assert np.abs(out[0][16:-18, 40:200] - true[16:-18, 40:200]).max() < 5e-3     # interior: away from the tapered margins
        assert tapered[0, 0].max() == 0.0 and np.array_equal(tapered[0, 20, 40:200], a[0, 20, 40:200])    # A is tapered in place


def small_db(nx=40, nz=6, ng=8):
    return Gfdb.create(nx, nz, ng, 0.1, 400.0, 400.0, 4000.0, 2000.0).build_ahfull(2700.0, 6000.0, 3464.0, nfflag=(ng == 10))


def test_oracle_database_interpolation_layout():
    db = small_db()
    meta, tr = ol.gfdb_interpolate(db, 2, 1)
    m = db.meta()
    assert meta["nx"] == 2 * m["nx"] and meta["nz"] == m["nz"] and abs(meta["dx"] - 200.0) < 1e-6
    s0, ln, off, dat = db.view()
    for ix in (1, 7, 40):
        for iz in (1, 6):
            for ig in (1, 8):
                k = ((ix - 1) * m["nz"] + iz - 1) * m["ng"] + ig - 1
                o0, v = tr[(2 * (ix - 1) + 1, iz, ig)]
                assert o0 == s0[k] and np.array_equal(v, dat[off[k]:off[k] + ln[k]])        # real traces are kept
    # an interpolated trace spans the union of its real neighbours (gfdb.f90:1199-1221); the last one repeats the end trace
    for ix in (2, 40, 80):
        a0, a = tr[(ix - 1, 3, 2)]
        b0, b = tr[(min(ix + 1, 79), 3, 2)]
        c0, c = tr[(ix, 3, 2)]
        assert c0 == min(a0, b0) and c0 + c.size == max(a0 + a.size, b0 + b.size)


@pytest.mark.gpu
@pytest.mark.parametrize("l1,l2,s1,s2,m1,m2", [(2, 1, 64, 1, 16, 0), (4, 1, 32, 1, 16, 0), (2, 2, 16, 32, 4, 16), (4, 4, 8, 16, 4, 16)])
def test_gulunay_operator_bit_exact(l1, l2, s1, s2, m1, m2):
    """kiwi_gulunay against the oracle's gulunay on random + plane-wave fields, several fields per call"""
    from kiwi_b200 import engine
    rng = np.random.default_rng(5)
    t = 128
    a = rng.standard_normal((3, s2, s1, t)).astype(np.float32)
    tt = np.arange(t)[None, None, :]
    a[1] = np.exp(-0.5 * ((tt - 40 - 0.7 * np.arange(s1)[None, :, None] - 0.3 * np.arange(s2)[:, None, None]) / 5.0) ** 2)
    a[2, :, :, 90:] = 0.0
    ag, og = engine.gulunay(a, l1, l2, 12, m1, m2)
    for b in range(3):
        ao, oo = ol.gulunay(a[b], l1, l2, 12, m1, m2)
        assert np.array_equal(ag[b].view(np.uint32), ao.view(np.uint32))
        assert np.array_equal(og[b].view(np.uint32), oo.view(np.uint32)), np.abs(og[b] - oo).max() / np.abs(oo).max()


@pytest.mark.gpu
@pytest.mark.parametrize("nipx,nipz,ng", [(2, 1, 10), (1, 2, 8), (2, 2, 8), (4, 2, 8), (4, 4, 8), (8, 1, 8), (1, 4, 8), (16, 1, 8)])
def test_database_interpolation_against_the_oracle(nipx, nipz, ng):
    db = small_db(36, 5, ng)
    meta, tr = ol.gfdb_interpolate(db, nipx, nipz)
    g = db.interpolate(nipx, nipz)
    m = g.meta()
    assert (m["nx"], m["nz"]) == (meta["nx"], meta["nz"]) and m["dx"] == np.float32(meta["dx"]) and m["dz"] == np.float32(meta["dz"])
    s0, ln, off, dat = g.view()
    assert int((ln > 0).sum()) == len(tr) == m["nx"] * m["nz"] * ng
    worst = 0.0
    for (ix, iz, ig), (o0, v) in tr.items():
        k = ((ix - 1) * m["nz"] + iz - 1) * ng + ig - 1
        assert s0[k] == o0 and ln[k] == v.size, (ix, iz, ig)
        if not np.array_equal(dat[off[k]:off[k] + ln[k]].view(np.uint32), v.view(np.uint32)):
            worst = max(worst, float(np.abs(dat[off[k]:off[k] + ln[k]] - v).max() / max(np.abs(v).max(), 1e-30)))
    assert worst == 0.0, worst


@pytest.mark.gpu
def test_engine_runs_on_an_interpolated_database():
    """set_database dbpath 2 1: synthetics on the interpolated grid stay close to those on an analytical grid of the same spacing"""
    from kiwi_b200 import Engine, synthetic
    coarse = Gfdb.create(60, 6, 8, 0.1, 400.0, 400.0, 4000.0, 2000.0).build_ahfull(2700.0, 6000.0, 3464.0, nfflag=False)
    fine = Gfdb.create(120, 6, 8, 0.1, 200.0, 400.0, 4000.0, 2000.0).build_ahfull(2700.0, 6000.0, 3464.0, nfflag=False)
    lat, lon, dep = synthetic.receivers(4, (30.0, 70.0), 9e3, 20e3, 3)
    p = np.array([0.0, 100, -200, 3000, 1e18, -0.4e18, -0.6e18, 0.3e18, 0.2e18, -0.5e18, 0.5], np.float32)
    out = []
    for db, nip in ((coarse, 2), (fine, 1)):
        e = Engine(0)
        e.set_database(db, nip, 1); e.set_local_interpolation("nearest_neighbor"); e.set_receivers(lat, lon, dep, ["ned"] * 4)
        e.set_source_location(30.0, 70.0, 0.0); e.set_effective_dt(0.5)
        e.set_source_params("moment_tensor", p)
        out.append([e.get_seismogram(ir, ic, 0) for ir in range(1, 5) for ic in range(1, 4)])
    for (fa, da), (fb, dbb) in zip(*out):
        lo, hi = max(fa, fb), min(fa + da.size, fb + dbb.size)
        x, y = da[lo - fa:hi - fa], dbb[lo - fb:hi - fb]
        assert np.dot(x, y) / np.sqrt(np.dot(x, x) * np.dot(y, y)) > 0.95


@pytest.mark.gpu
def test_interpolation_refuses_what_the_reference_cannot_do():
    from kiwi_b200 import KiwiError, engine
    db = small_db(12, 2, 8)
    with pytest.raises(KiwiError, match="interpolation factors must be"):
        db.interpolate(3, 1)
    holed = Gfdb.create(12, 2, 8, 0.1, 400.0, 400.0, 4000.0, 2000.0)
    holed.save_array(1, 1, 1, 10, np.ones(8, np.float32))              # one trace only
    with pytest.raises(KiwiError, match="[Mm]issing trace"):
        holed.interpolate(2, 1)
    with pytest.raises(KiwiError, match="powers of two"):
        engine.gulunay(np.zeros((1, 1, 12, 100), np.float32), 2, 1, 8, 4, 0)
    with pytest.raises(KiwiError, match="margins overlap"):
        engine.gulunay(np.zeros((1, 1, 4, 64), np.float32), 2, 1, 8, 16, 0)
