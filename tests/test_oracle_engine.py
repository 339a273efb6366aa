"""Properties of the oracle itself on the unpinned parts of the path (SURVEY.md section 8c:
make_seismogram, bilinear fetch, rotations, scaling have no reference test) -- cross-checks that do
not depend on the CUDA path."""
import numpy as np

import scenario as sc
from oracle_lib import OracleEngine

COMPS = ["ned", "ar", "d", "neu", "cl", "wsd"]


def make(comps=COMPS, **kw):
    lat, lon, dep = sc.small_receivers(len(comps))
    o = OracleEngine()
    sc.setup(o, sc.small_db(), lat, lon, dep, comps, **kw)
    return o


def seis(o, stype, p, ir, ic, which=0):
    o.eval_sources(stype, p)
    return o.get_seismogram(ir, ic, which)


def test_linearity_in_moment_tensor():
    o = make()
    a = sc.MT_SMALL.copy(); b = sc.MT_SMALL.copy(); c = sc.MT_SMALL.copy()
    b[4:10] = [0.3e18, 0.1e18, -0.2e18, 0.5e18, -0.1e18, 0.25e18]
    c[4:10] = a[4:10] + b[4:10]
    for ir, ic in ((1, 1), (1, 3), (2, 2), (4, 1)):
        fa, da = seis(o, "moment_tensor", a, ir, ic)
        fb, db_ = seis(o, "moment_tensor", b, ir, ic)
        fc, dc = seis(o, "moment_tensor", c, ir, ic)
        assert fa == fb == fc
        assert np.abs(da + db_ - dc).max() <= 5e-6 * np.abs(dc).max()


def test_integer_time_shift_moves_samples():
    o = make()
    a = sc.MT_SMALL.copy(); b = a.copy(); b[0] += 0.5   # 5 samples at dt = 0.1
    fa, da = seis(o, "moment_tensor", a, 3, 1)
    fb, db_ = seis(o, "moment_tensor", b, 3, 1)
    assert fb - fa == 5 and da.size == db_.size
    assert np.abs(da - db_).max() <= 2e-5 * np.abs(da).max()


def test_opposite_components_are_negatives():
    o = make(["ned", "swu", "ar", "cl"])
    # receivers 1/2 and 3/4 differ in position, so compare within one receiver via two setups
    lat, lon, dep = sc.small_receivers(1)
    o1, o2 = OracleEngine(), OracleEngine()
    sc.setup(o1, sc.small_db(), lat, lon, dep, ["nedar"])
    sc.setup(o2, sc.small_db(), lat, lon, dep, ["swucl"])
    o1.eval_sources("bilateral", sc.BILAT_SMALL); o2.eval_sources("bilateral", sc.BILAT_SMALL)
    for ic in range(1, 6):
        f1, d1 = o1.get_seismogram(1, ic); f2, d2 = o2.get_seismogram(1, ic)
        assert f1 == f2 and np.array_equal(d1, -d2)


def test_north_east_is_rotated_away_right():
    lat, lon, dep = sc.small_receivers(1)
    o = OracleEngine()
    sc.setup(o, sc.small_db(), lat, lon, dep, ["nedar"])
    o.eval_sources("moment_tensor", sc.MT_SMALL)
    fn, n = o.get_seismogram(1, 1); fe, e = o.get_seismogram(1, 2); fa, a = o.get_seismogram(1, 4); fr, r = o.get_seismogram(1, 5)
    lo = max(fn, fa, fr); hi = min(fn + n.size, fa + a.size, fr + r.size)
    sl = lambda f, d: d[lo - f:hi - f]
    # rotation preserves the horizontal vector length (seismogram.f90:268-283)
    h1 = sl(fn, n) ** 2 + sl(fe, e) ** 2
    h2 = sl(fa, a) ** 2 + sl(fr, r) ** 2
    assert np.abs(h1 - h2).max() <= 1e-5 * h2.max()


def test_misfit_of_scaled_reference_and_global_formula():
    o = make()
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [o], [len(c) for c in COMPS], scale=1.07)
    for norm, expect in (("l2norm", 0.07 / 1.07), ("l1norm", 0.07 / 1.07)):
        o.set_misfit_method(norm)
        m, st = o.eval_sources("bilateral", sc.BILAT_SMALL)
        assert not st.any()
        ratio = m[0, :, 0] / m[0, :, 1]
        assert np.allclose(ratio, expect, rtol=2e-4)
        gm = np.sqrt((m[0, :, 0].astype(np.float64) ** 2).sum()) / np.sqrt((m[0, :, 1].astype(np.float64) ** 2).sum())
        assert abs(o.get_global_misfit() - gm) <= 1e-5 * gm   # minimizer_engine.f90:939-942


def test_moment_only_scales_bilateral_synthetics():
    o = make()
    a = sc.BILAT_SMALL.copy(); b = a.copy(); b[4] *= 2.5
    fa, da = seis(o, "bilateral", a, 1, 1, which=1)
    fb, db_ = seis(o, "bilateral", b, 1, 1, which=1)
    assert fa == fb and np.allclose(db_, 2.5 * da, rtol=1e-6)


def test_fresh_state_makes_evaluation_order_irrelevant():
    o = make()
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [o], [len(c) for c in COMPS])
    p2 = sc.BILAT_SMALL.copy(); p2[0] += 2.0; p2[9] += 1500
    m1, _ = o.eval_sources("bilateral", np.stack([sc.BILAT_SMALL, p2]))
    m2, _ = o.eval_sources("bilateral", np.stack([p2, sc.BILAT_SMALL]))
    assert np.array_equal(m1[0], m2[1]) and np.array_equal(m1[1], m2[0])


def test_eikonal_source_properties():
    """source_eikonal.f90: weights sum to one, times are centred (centertime), points respect the default
    depth constraints (parameterized_source.f90:127-145), the rise-time fold keeps the static offset."""
    o = make()
    p = np.array([0.1, 100, -200, 3500, 2e18, 40, 70, 20, 0, 0, 1500, 300, -200, 0.8, 0.4], np.float32)
    t, g, n = o.discretize_source("eikonal", p)
    assert n > 50 and g[0] >= 2 and g[1] >= 2
    m = t[:, 4:10].astype(np.float64)
    # sum of the centroid tensors = unit double couple (weights and time weights both sum to one)
    assert abs(np.abs(m.sum(0)).max() - 1.0) < 0.05 or np.linalg.norm(m.sum(0)) > 0.9
    assert t[:, 2].min() >= 1500.0
    assert np.hypot(t[:, 0] - 100, t[:, 1] + 200).max() <= 1500.0 * 1.01
    o.eval_sources("eikonal", p)
    f1, folded = o.get_seismogram(1, 3, 1)
    f0, raw = o.get_seismogram(1, 3, 0)
    assert abs(folded[-1] - raw[-1] * 2e18) <= 1e-4 * abs(folded).max()      # boxcar weights sum to one
    p2 = p.copy(); p2[11] = 9000
    _, st = o.eval_sources("eikonal", p2)
    assert st[0] == 1


def test_autoshift_finds_the_delay_of_the_references():
    """receiver_autoshift_ref_seismogram (receiver.f90:816-832) has no reference test: references that are the synthetics delayed
    by k samples must be moved back by k, after which the misfit is that of the unshifted references"""
    ncomps = [len(c) for c in COMPS]
    lat, lon, dep = sc.small_receivers(len(COMPS))
    a, b = OracleEngine(), OracleEngine()
    for o in (a, b):   # far-field terms only: the traces return to zero, so the cross-correlation peaks at the true delay
        sc.setup(o, sc.small_db_ng8(), lat, lon, dep, COMPS)
    a.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(a, [a], ncomps, shift=0)
    sc.set_refs_from(a, [b], ncomps, shift=4)
    m0, _ = a.eval_sources("bilateral", sc.BILAT_SMALL)
    m_late, _ = b.eval_sources("bilateral", sc.BILAT_SMALL)
    assert m_late[0, :, 0].sum() > 2 * m0[0, :, 0].sum()
    b.shift_ref_seismogram(6, -0.1)                      # receiver 6: late by 3 only
    shifts = b.autoshift_ref_seismogram(0, -0.7, 0.7)
    assert np.allclose(shifts, [-0.4] * 5 + [-0.3], atol=1e-6)
    m1, _ = b.eval_sources("bilateral", sc.BILAT_SMALL)
    assert np.array_equal(m0, m1)
    # a second pass finds nothing left to do
    assert np.allclose(b.autoshift_ref_seismogram(0, -0.7, 0.7), 0.0)
