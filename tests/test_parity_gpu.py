"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for discretisation, GF indices, sample shifts and spans; 1e-5 relative (north_star) for
seismograms and misfits."""
import numpy as np
import pytest

import scenario as sc
from oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # BASELINE.json north_star: "within 1e-5 relative for seismograms and misfits (fp32)"


def misfit_tol(mo, floor=0.1):
    """Tolerance for a block [..., 2] of (misfit, norm factor) pairs: 1e-5 relative.  A misfit far
    below its norm factor is a difference of nearly equal fp32 traces (e.g. the +7 % reference of
    SURVEY.md 8d evaluated at the true source: m = 0.065 nf); the traces themselves only agree to
    1e-5 of their size, so below m = 0.1 nf the bound is relative to 0.1 nf instead of m."""
    return RTOL * np.maximum(np.abs(mo), floor * np.abs(mo[..., 1:2]))

COMPS6 = ["ned", "ar", "d", "neu", "cl", "wsd"]


def engines(db, comps, n=6, **kw):
    from kiwi_b200 import Engine
    lat, lon, dep = sc.small_receivers(n)
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, **kw)
    return g, o


def assert_seis_close(a, b, label):
    (fa, da), (fb, dbb) = a, b
    assert fa == fb and da.size == dbb.size, "%s: span differs: gpu [%d,+%d) oracle [%d,+%d)" % (label, fa, da.size, fb, dbb.size)
    scale = np.abs(dbb).max()
    assert scale > 0
    err = np.abs(da - dbb).max() / scale
    assert err <= RTOL, "%s: max deviation %.3g of peak" % (label, err)


@pytest.mark.parametrize("params", [sc.BILAT_SMALL,
                                    np.array([0, 0, 0, 2500, 1e18, 10, 45, -90, 0, 1500, 1500, 3000, 2500, 0.0], np.float32),
                                    np.array([1, 100, 100, 2000, 1e18, 200, 89, 10, 90, 0, 0, 0, 3000, 1.0], np.float32)])
def test_bilateral_discretisation_bit_exact(params):
    g, o = engines(sc.small_db(), COMPS6)
    tg, gg, ng = g.discretize_source("bilateral", params)
    to, go, no = o.discretize_source("bilateral", params)
    assert ng == no and list(gg) == list(go)
    assert np.array_equal(tg.view(np.uint32), to.view(np.uint32))


def test_moment_tensor_discretisation_bit_exact():
    g, o = engines(sc.small_db(), COMPS6)
    tg, gg, ng = g.discretize_source("moment_tensor", sc.MT_SMALL)
    to, go, no = o.discretize_source("moment_tensor", sc.MT_SMALL)
    assert ng == no == 4 and gg[2] == go[0]
    assert np.array_equal(tg.view(np.uint32), to.view(np.uint32))


@pytest.mark.parametrize("interp,under", [("bilinear", (1, 1)), ("nearest_neighbor", (1, 1)), ("bilinear", (2, 3))])
def test_indices_shifts_spans_bit_exact(interp, under):
    g, o = engines(sc.small_db(), COMPS6, interpolation=interp, under=under)
    o.record_indices(True)
    o.eval_sources("bilateral", sc.BILAT_SMALL)   # no references yet: status 1, seismograms are there
    g.set_source_params("bilateral", sc.BILAT_SMALL)
    nflag = 0
    for ir in range(1, 7):
        ig, io = g.get_indices(ir), o.get_indices(ir)
        assert ig["ix"].size == io["ix"].size == 204
        ok = ig["near"] == 0   # the device still reports the pairs whose coordinate sits on a cell edge
        nflag += int((~ok).sum())
        # all of them equal, flagged or not, to the last bit: since round 2 the one transcendental of the distance chain whose device
        # version differs from glibc's often enough to matter (atan2f of the sub-source azimuth) comes from the host library
        for k in ("ix", "iz", "its"):
            assert np.array_equal(ig[k], io[k]), k
        assert np.array_equal(ig["dix"].view(np.uint32), io["dix"].view(np.uint32))
        assert np.array_equal(ig["diz"].view(np.uint32), io["diz"].view(np.uint32))
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            fg, dg = g.get_seismogram(ir, ic)
            fo, do = o.get_seismogram(ir, ic)
            assert (fg, dg.size) == (fo, do.size), "span of receiver %d component %d" % (ir, ic)
    assert nflag < 10


@pytest.mark.parametrize("stype,params", [("bilateral", sc.BILAT_SMALL), ("moment_tensor", sc.MT_SMALL)])
@pytest.mark.parametrize("interp", ["bilinear", "nearest_neighbor"])
def test_seismograms(stype, params, interp):
    g, o = engines(sc.small_db(), COMPS6, interpolation=interp)
    o.eval_sources(stype, params)
    g.set_source_params(stype, params)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            for which in (0, 1):
                assert_seis_close(g.get_seismogram(ir, ic, which), o.get_seismogram(ir, ic, which), "rcv %d comp %d which %d" % (ir, ic, which))


def test_seismograms_ng8():
    g, o = engines(sc.small_db_ng8(), COMPS6)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    g.set_source_params("bilateral", sc.BILAT_SMALL)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            assert_seis_close(g.get_seismogram(ir, ic), o.get_seismogram(ir, ic), "rcv %d comp %d" % (ir, ic))


def _candidates():
    p = np.tile(sc.BILAT_SMALL, (7, 1))
    p[1, 5] += 15; p[2, 6] -= 20; p[3, 7] += 40; p[4, 3] += 500; p[5, 9] += 800; p[6, 4] *= 1.3
    p[6, 13] = 0.2
    return p


@pytest.mark.parametrize("norm", ["l2norm", "l1norm", "scalar_product", "peak"])
@pytest.mark.parametrize("taper", [False, True])
def test_misfits_batched(norm, taper):
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps)
    for e in (g, o):
        e.set_misfit_method(norm)
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, [1.0, 1.6, 4.0, 5.2], [0, 1, 1, 0])
    p = _candidates()
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    assert np.array_equal(sg, so) and not sg.any()
    assert mg.shape == mo.shape == (7, 14, 2)
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()
    # ns = 1 pair with the reference's names gives the same numbers as the batch
    g.set_source_params("bilateral", p[3])
    assert np.array_equal(g.get_misfits(), mg[3])
    from kiwi_b200 import global_misfits
    gm = global_misfits(mg)
    for i in (0, 3):
        o.eval_sources("bilateral", p[i])
        assert abs(gm[i] - o.get_global_misfit()) <= 1e-5 * max(abs(o.get_global_misfit()), 1e-3)


def test_disabled_receivers_and_moment_tensor_batch():
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("moment_tensor", sc.MT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps)
    for e in (g, o):
        e.switch_receiver(2, False); e.switch_receiver(5, False)
        e.set_misfit_method("l1norm")
    assert g.nmisfits == o.nmisfits == 10
    p = np.tile(sc.MT_SMALL, (5, 1))
    p[1, 4:10] *= -0.5; p[2, 1] += 400; p[3, 3] -= 600; p[4, 10] = 0.25
    mg, sg = g.eval_sources("moment_tensor", p)
    mo, so = o.eval_sources("moment_tensor", p)
    assert not sg.any() and not so.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo))


def test_error_behaviour():
    from kiwi_b200 import Engine, KiwiError
    e = Engine(0)
    with pytest.raises(KiwiError, match="no database set"):
        e.set_receivers([30.1], [70.1], [0], ["ned"])
    e.set_database(sc.small_db())
    with pytest.raises(KiwiError, match="no receivers set"):
        e.eval_sources("bilateral", sc.BILAT_SMALL)
    lat, lon, dep = sc.small_receivers(2)
    with pytest.raises(KiwiError, match="initializing receiver failed"):
        e.set_receivers(lat, lon, dep, ["ns", "d"])
    e.set_receivers(lat, lon, dep, ["ned", "d"])
    with pytest.raises(KiwiError, match="no source location set"):
        e.eval_sources("bilateral", sc.BILAT_SMALL)
    e.set_source_location(*sc.ORIGIN, 0.0)
    with pytest.raises(KiwiError, match="no reference seismograms set"):
        e.eval_sources("bilateral", sc.BILAT_SMALL)
    with pytest.raises(KiwiError, match="no source parameters set"):
        e.get_misfits()
    with pytest.raises(KiwiError, match="wrong number"):
        e.eval_sources("bilateral", sc.BILAT_SMALL[:5])


def test_medium_source_against_fp32_and_double_accumulating_oracle():
    """~2000 sub-sources: the fp32 reference path's sequential accumulation noise grows with the number
    of centroids, so next to the fp32 restatement the CUDA result is also held against the
    restatement with strips carried in double (oracle -DKO_WIDE)."""
    from kiwi_b200 import Engine
    comps = ["ned"] * 4
    lat, lon, dep = sc.small_receivers(4)
    p = sc.BILAT_SMALL.copy(); p[9] = 4000; p[10] = 3000; p[11] = 3000
    g, o, ow = Engine(0), OracleEngine(), OracleEngine(wide=True)
    for e in (g, o, ow):
        sc.setup(e, sc.small_db(), lat, lon, dep, comps, effective_dt=0.05)
    o.eval_sources("bilateral", p)
    assert o.discretize_source("bilateral", p)[2] > 1500
    sc.set_refs_from(o, [g, o, ow], [3] * 4)
    cands = np.tile(p, (3, 1)); cands[1, 5] += 12; cands[2, 3] += 300
    mg, _ = g.eval_sources("bilateral", cands)
    mo, _ = o.eval_sources("bilateral", cands)
    mw, _ = ow.eval_sources("bilateral", cands)
    dev = lambda a, b: float(np.max(np.abs(a - b) / (misfit_tol(b) / RTOL)))
    d32, dw, d32w = dev(mg, mo), dev(mg, mw), dev(mo, mw)
    assert dw <= RTOL, (d32, dw, d32w)
    assert d32 <= max(RTOL, 2.0 * d32w), (d32, dw, d32w)


FILTER = ([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])      # band-pass, Hz (shape of python/tunguska/filtering.py:14-19)
TAPER = ([1.0, 1.6, 4.0, 5.2], [0, 1, 1, 0])


def _spectral_setup(norm, taper, filt, comps=COMPS6, wide=False):
    g, o = engines(sc.small_db(), comps)
    ncomps = [len(c) for c in comps]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    es = [g, o]
    if wide:   # the restatement with strips carried in double: arbiter where the fp32 path's own rounding noise is near the bar
        lat, lon, dep = sc.small_receivers(len(comps))
        w = OracleEngine(wide=True)
        sc.setup(w, sc.small_db(), lat, lon, dep, comps)
        es.append(w)
    sc.set_refs_from(o, es, ncomps)
    for e in es:
        e.set_misfit_method(norm)
        if taper:
            for ir in range(1, len(comps) + 1):
                e.set_misfit_taper(ir, *TAPER)
        if filt:
            e.set_misfit_filter(*FILTER)
    return es


@pytest.mark.parametrize("norm", ["ampspec_l2norm", "ampspec_l1norm"])
@pytest.mark.parametrize("taper,filt", [(False, False), (True, False), (True, True), (False, True)])
def test_amplitude_spectrum_misfits(norm, taper, filt):
    """comparator.f90:861-909 through the shared-memory FFT kernel.  The reference's FFT is FFTW
    (unpinned third-party arithmetic); two correct fp32 FFTs differ by ~1e-6 of the
    spectral peak per bin, so for amplitude-spectrum misfits far below their norm factor the 1e-5
    bound is taken relative to 0.25 nf (0.1 nf for time-domain norms)."""
    g, o = _spectral_setup(norm, taper, filt)
    p = _candidates()
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    assert not sg.any() and not so.any()
    tol = misfit_tol(mo, 0.25)
    assert np.all(np.abs(mg - mo) <= tol), np.abs((mg - mo) / tol).max()


@pytest.mark.parametrize("norm", ["l2norm", "l1norm", "scalar_product", "peak"])
@pytest.mark.parametrize("taper", [False, True])
def test_filtered_time_domain_misfits(norm, taper):
    """make_spectrum_filtered / make_array_filtered (comparator.f90:1217-1263): r2c -> filter -> c2r -> norm"""
    g, o, w = _spectral_setup(norm, taper, True, wide=True)
    p = _candidates()
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    mw, sw = w.eval_sources("bilateral", p)
    assert not sg.any() and not so.any() and not sw.any()
    # candidate 0 is the source the references were made from (x 1.07): m = 0.065 nf is a difference of nearly equal
    # traces, and without a taper the filter spreads every rounding of the synthetic over the whole padded span.  The fp32
    # restatement itself sits at ~1e-5 of its double-accumulated twin there, so the bar is 1e-5 against the latter and
    # twice the fp32 path's own distance against the former.
    tol = misfit_tol(mw)
    assert np.all(np.abs(mg - mw) <= tol), np.abs((mg - mw) / tol).max()
    assert np.all(np.abs(mg - mo) <= np.maximum(tol, 2.0 * np.abs(mo - mw))), np.abs((mg - mo) / tol).max()


@pytest.mark.parametrize("norm", ["floating_l2norm", "floating_l1norm"])
@pytest.mark.parametrize("taper,filt", [(False, False), (True, False), (True, True)])
def test_floating_misfits(norm, taper, filt):
    """receiver_calculate_floating_misfits (receiver.f90:439-510): best integer shift over all components"""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps, shift=3)          # references late by 3 samples
    for e in (g, o):
        e.set_misfit_method(norm)
        e.set_floating_shiftrange(-0.6, 0.5)
        e.set_floating_shiftrange(-0.2, 0.9, 2)      # receiver 2: a range that excludes the true shift
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
        if filt:
            e.set_misfit_filter(*FILTER)
    p = _candidates()
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    assert not sg.any() and not so.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()
    # the shift itself is an integer result: exact
    o.eval_sources("bilateral", p[0])
    g.set_source_params("bilateral", p[0])
    g.get_misfits()
    assert list(g.get_floating_shifts()) == list(o.get_floating_shifts())
    assert list(g.get_floating_shifts())[0] == -3


@pytest.mark.parametrize("premethod", ["l2norm", "floating_l2norm"])
@pytest.mark.parametrize("taper,filt", [(False, False), (True, False), (True, True)])
def test_autoshift_ref_seismogram(premethod, taper, filt):
    """autoshift_ref_seismogram (minimizer_engine.f90:380-416, receiver.f90:816-832): the references move to the maximum of the
    windowed cross-correlation with the synthetics; shift_ref_seismogram (:354-378) moves them by hand.  Far-field database:
    with the static near-field offsets the continuation rule makes the correlation double-peaked with an exact tie."""
    g, o = engines(sc.small_db_ng8(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps, shift=3)          # references late by 3 samples
    for e in (g, o):
        e.set_misfit_method(premethod)
        if premethod.startswith("floating"):
            e.set_floating_shiftrange(-0.2, 0.2)
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
        if filt:
            e.set_misfit_filter(*FILTER)
        e.switch_receiver(4, False)
        e.shift_ref_seismogram(5, 0.2)                    # receiver 5: late by 5 samples now
    p = _candidates()
    o.eval_sources("bilateral", p[0])
    g.set_source_params("bilateral", p[0])
    sg = g.autoshift_ref_seismogram(2, -0.1, 0.9)          # one receiver, a range that excludes the true shift
    so = o.autoshift_ref_seismogram(2, -0.1, 0.9)
    assert np.array_equal(sg, so) and sg.size == 1
    sg = g.autoshift_ref_seismogram(0, -0.8, 0.6)
    so = o.autoshift_ref_seismogram(0, -0.8, 0.6)
    assert np.array_equal(sg, so) and sg.size == 6
    assert sg[3] == 0.0                                   # disabled receiver
    if not taper:
        assert abs(sg[0] + 0.3) < 1e-6 and abs(sg[4] + 0.5) < 1e-6
    # the shifted references are what the next evaluations see
    mg, stg = g.eval_sources("bilateral", p)
    mo, sto = o.eval_sources("bilateral", p)
    assert not stg.any() and not sto.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()
    with pytest.raises(Exception, match="receiver index out of range"):
        g.autoshift_ref_seismogram(7, -0.1, 0.1)


@pytest.mark.parametrize("taper,filt", [(False, False), (True, False), (True, True)])
def test_cross_correlations(taper, filt):
    """receiver_calculate_cross_correlations (receiver.f90:597-616, comparator.f90:1061-1090) as output_cross_correlations returns them"""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps, shift=-2)
    for e in (g, o):
        e.set_synthetics_factor(0.9)
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
        if filt:
            e.set_misfit_filter(*FILTER)
    p = _candidates()[1]
    o.eval_sources("bilateral", p)
    g.set_source_params("bilateral", p)
    for ir in range(1, 7):
        cg = g.get_cross_correlations(ir, -0.5, 0.7)
        co = o.get_cross_correlations(ir, -0.5, 0.7)
        assert cg.shape == co.shape == (ncomps[ir - 1], 13)
        scale = np.abs(co).max(axis=1, keepdims=True)
        assert np.all(np.abs(cg - co) <= (4 if filt else 1) * RTOL * scale), (ir, (np.abs(cg - co) / scale).max())


@pytest.mark.parametrize("taper,filt", [(False, False), (True, False), (False, True), (True, True)])
def test_probe_export_all_variants(taper, filt):
    """output_seismograms / output_seismogram_spectra in memory (probe_get, probe_get_amp_spectrum comparator.f90:332-433): synthetics and
    references, plain / tapered / filtered, with the spans of a fresh process"""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps, shift=2)
    for e in (g, o):
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
        if filt:
            e.set_misfit_filter(*FILTER)
    p = _candidates()[6]                     # with a rise time: the folded synthetic
    o.eval_sources("bilateral", p)
    g.set_source_params("bilateral", p)
    for ir, ic in ((1, 1), (2, 2), (3, 1), (5, 2), (6, 3)):
        for which in ("synthetics", "references"):
            for proc in ("plain", "tapered", "filtered"):
                fg, dg = g.get_probe(ir, ic, which, proc)
                fo, do = o.get_probe(ir, ic, which, proc)
                assert fg == fo and dg.size == do.size, (ir, ic, which, proc, fg, fo, dg.size, do.size)
                scale = np.abs(do).max()
                lim = (4 if (filt and proc == "filtered") else 1) * RTOL * scale
                assert np.abs(dg - do).max() <= lim, (ir, ic, which, proc, np.abs(dg - do).max() / scale)
                if not (filt and proc == "filtered") and which == "references":
                    assert np.array_equal(dg, do)          # no arithmetic but one multiplication
                dfg, ag = g.get_probe(ir, ic, which, proc, spectrum=True)
                dfo, ao = o.get_probe(ir, ic, which, proc, spectrum=True)
                assert dfg == dfo and ag.size == ao.size
                assert np.abs(ag - ao).max() <= 4 * RTOL * np.abs(ao).max(), (ir, ic, which, proc, np.abs(ag - ao).max() / np.abs(ao).max())
    with pytest.raises(Exception, match="component index out of range"):
        g.get_probe(3, 2)
    with pytest.raises(Exception, match="receiver index out of range"):
        g.get_cross_correlations(0, -0.1, 0.1)
    with pytest.raises(Exception, match="empty shift range"):
        g.get_cross_correlations(1, 0.3, 0.1)
    fresh, _ = engines(sc.small_db(), COMPS6)
    fresh.set_source_params("bilateral", p)
    with pytest.raises(Exception, match="no reference seismograms set"):
        fresh.get_probe(1, 1, "references")
    with pytest.raises(Exception, match="no reference seismograms set"):
        fresh.autoshift_ref_seismogram(0, -0.1, 0.1)
    assert fresh.get_probe(1, 1, "synthetics")[1].size > 10


def test_distances_crustal_thickness_and_principal_axes():
    """get_distances (minimizer_engine.f90:1260-1281), get_source_crustal_thickness (:488-498), get_principal_axes (:1248-1258): host
    arithmetic of the reference, exact against the restatement"""
    g, o = engines(sc.small_db(), COMPS6)
    dg, ag = g.get_distances()
    do, ao = o.get_distances()
    assert dg.size == 6 and np.array_equal(dg, do) and np.array_equal(ag, ao)
    assert g.get_source_crustal_thickness() == o.get_source_crustal_thickness() > 5000.0
    for e in (g, o):
        e.set_source_crustal_thickness_limit(12345.0)
    assert g.get_source_crustal_thickness() == o.get_source_crustal_thickness() == 12345.0
    for stype, p in (("bilateral", sc.BILAT_SMALL), ("circular", CIRC), ("eikonal", EIK)):
        g.set_source_params(stype, p)
        pg, tg = g.get_principal_axes()
        po, to = o.principal_axes(float(p[5]), float(p[6]), float(p[7]))
        assert np.array_equal(pg, po) and np.array_equal(tg, to) and np.all(np.abs(pg) <= 180.0)
    g.set_source_params("moment_tensor", sc.MT_SMALL)
    pg, tg = g.get_principal_axes()
    assert not pg.any() and not tg.any()


def _mt_grid(nloc=3, nmt=45):
    from kiwi_b200 import synthetic
    mts = synthetic.fibonacci_moment_tensors(nmt) * 1e18
    rng = np.random.default_rng(3)
    mts = mts * rng.uniform(0.3, 2.0, (nmt, 1)).astype(np.float32)
    p = np.zeros((nloc, nmt, 11), np.float32)
    for l in range(nloc):
        p[l, :, 0] = 0.1 * l; p[l, :, 1] = 150 + 300 * l; p[l, :, 2] = -250 + 200 * l; p[l, :, 3] = 2800 - 400 * l
        p[l, :, 4:10] = mts; p[l, :, 10] = 0.7
    p = p.reshape(-1, 11)
    return p[rng.permutation(p.shape[0])]      # locations interleaved: the engine has to find the grid itself


@pytest.mark.parametrize("norm", ["l2norm", "l1norm"])
@pytest.mark.parametrize("taper", [False, True])
def test_moment_tensor_grid_search_tensor_core_path(norm, taper):
    """config C2: candidates sharing a location are contracted against six unit-tensor basis synthetics
    with tcgen05 (3xTF32); held against the oracle and against the engine's own direct path"""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("moment_tensor", sc.MT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps)
    for e in (g, o):
        e.set_misfit_method(norm)
        if taper:
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
    p = _mt_grid()
    mg, sg = g.eval_sources("moment_tensor", p)
    assert g.last_timing()["launches"][3] >= 1
    mo, so = o.eval_sources("moment_tensor", p)
    assert not sg.any() and not so.any()
    assert np.all(np.abs(mg - mo) <= misfit_tol(mo)), np.abs((mg - mo) / misfit_tol(mo)).max()
    g.set_mt_grid(2)                       # the unfused variant: basis seismograms through k_synth, then k_mt_contract
    mu, su = g.eval_sources("moment_tensor", p)
    assert g.last_timing()["launches"][2] >= 1 and g.last_timing()["launches"][3] >= 1 and not su.any()
    assert np.all(np.abs(mu - mo) <= misfit_tol(mo)), np.abs((mu - mo) / misfit_tol(mo)).max()
    g.set_mt_grid(False)
    md, sd = g.eval_sources("moment_tensor", p)
    assert np.all(np.abs(mg - md) <= misfit_tol(mo)), np.abs((mg - md) / misfit_tol(mo)).max()
    assert not np.array_equal(mg, md) and not np.array_equal(mg, mu) and not np.array_equal(mu, md)     # really three different code paths


# eikonal: time north east depth moment strike dip rake bord-x bord-y bord-radius nukl-x nukl-y rel-rupture-velocity rise-time
EIK = np.array([0.1, 100, -200, 3500, 2e18, 40, 70, 20, 0, 0, 1500, 300, -200, 0.8, 0.4], np.float32)
# mt_eikonal: ... dip bord-x bord-y bord-radius nukl-x nukl-y rel-rupture-velocity mxx myy mzz mxy mxz myz rise-time
MTEIK = np.array([0.1, 100, -200, 3500, 1.5, 40, 70, 0, 0, 1500, 300, -200, 0.8, 1e18, -0.4e18, -0.6e18, 0.3e18, 0.2e18, -0.5e18, 0.3], np.float32)


@pytest.mark.parametrize("stype,params", [("eikonal", EIK), ("mt_eikonal", MTEIK),
                                          ("eikonal", np.array([0, 0, 0, 2600, 1e18, 120, 85, -60, 200, -300, 2500, -400, 100, 0.9, 0.0], np.float32))])
def test_eikonal_discretisation_bit_exact(stype, params):
    """source_eikonal.f90 / source_mt_eikonal.f90 incl. the fast-marching solve (eikonal.f90, heap.f90): the whole
    centroid table bit for bit"""
    g, o = engines(sc.small_db(), COMPS6)
    tg, gg, ng = g.discretize_source(stype, params)
    to, go, no = o.discretize_source(stype, params)
    assert ng == no and ng > 50
    assert list(gg[:2]) == list(go[:2])
    assert np.array_equal(tg.view(np.uint32), to.view(np.uint32))


@pytest.mark.parametrize("stype,params", [("eikonal", EIK), ("mt_eikonal", MTEIK)])
def test_eikonal_seismograms_with_rise_time_fold(stype, params):
    """synthesis + rise-time boxcar fold (receiver.f90:853-904, strip_fold) + moment scaling.
    strip_fold cuts the strip at strip_dataspan, whose end is "where the exactly constant tail begins"
    (sparse_trace.f90:366-374).  In the reference that tail is only constant up to fp32 accumulation noise
    (each tail sample sums the same end values in its own rounding), so where the *exactly* constant run
    starts is decided by that noise: the folded strip may be a few samples longer or shorter.  The strips
    are therefore compared on the union span with the continuation rule (last sample repeats); the
    first sample index is exact."""
    g, o = engines(sc.small_db(), COMPS6)
    o.eval_sources(stype, params)
    g.set_source_params(stype, params)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic, 1), o.get_seismogram(ir, ic, 1)
            assert fg == fo and abs(dg.size - do.size) <= 8, (ir, ic, fg, fo, dg.size, do.size)
            n = max(dg.size, do.size)
            eg = np.concatenate([dg, np.full(n - dg.size, dg[-1], np.float32)])
            eo = np.concatenate([do, np.full(n - do.size, do[-1], np.float32)])
            assert np.abs(eg - eo).max() <= RTOL * np.abs(eo).max(), (ir, ic, np.abs(eg - eo).max() / np.abs(eo).max())


@pytest.mark.parametrize("norm", ["l2norm", "ampspec_l1norm"])
def test_eikonal_misfits_batched_with_failures(norm):
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("eikonal", EIK)
    sc.set_refs_from(o, [g, o], ncomps)
    for e in (g, o):
        e.set_misfit_method(norm)
        for ir in range(1, 7):
            e.set_misfit_taper(ir, *TAPER)
        e.set_misfit_filter(*FILTER)
    p = np.tile(EIK, (6, 1))
    p[1, 5] += 25; p[2, 10] = 2200; p[3, 13] = 0.7; p[3, 14] = 0.8
    p[4, 11] = 5000          # nucleation point outside the rupture area: this candidate fails, the batch goes on
    p[5, 3] = 900; p[5, 10] = 300; p[5, 11] = 0; p[5, 12] = 0   # whole circle above the 1500 m plane: "Empty rupture area"
    mg, sg = g.eval_sources("eikonal", p)
    mo, so = o.eval_sources("eikonal", p)
    assert list(sg) == [0, 0, 0, 0, 1, 1] and list(so) == [0, 0, 0, 0, 1, 1]
    tol = misfit_tol(mo[:4], 0.25 if norm.startswith("ampspec") else 0.1)
    assert np.all(np.abs(mg[:4] - mo[:4]) <= tol), np.abs((mg[:4] - mo[:4]) / tol).max()
    assert np.all(np.isnan(mg[4:]))


def test_eikonal_user_constraints_and_thickness_limit():
    g, o = engines(sc.small_db(), COMPS6)
    for e in (g, o):
        e.set_source_crustal_thickness_limit(4000.0)
    tg, _, ng = g.discretize_source("eikonal", EIK)
    to, _, no = o.discretize_source("eikonal", EIK)
    assert ng == no and np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    assert tg[:, 2].max() <= 4000.0
    pts, nrm = [[0, 0, 2500], [0, 0, 4500], [0, -300, 0]], [[0, 0, -1], [0, 0, 1], [0, -1, 0]]
    for e in (g, o):
        e.set_source_constraints(pts, nrm)
    tg, _, ng = g.discretize_source("eikonal", EIK + np.float32(0))
    to, _, no = o.discretize_source("eikonal", EIK)
    assert ng == no and np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    assert tg[:, 2].min() >= 2500.0 and tg[:, 1].min() >= -300.0


@pytest.mark.parametrize("stype,base,norm", [("bilateral", sc.BILAT_SMALL, "l1norm"), ("eikonal", EIK, "floating_l2norm"), ("bilateral", sc.BILAT_SMALL, "ampspec_l2norm")])
def test_moment_axis_shares_syntheses(stype, base, norm):
    """A grid axis over the moment (python/examples/kiwi: moment x0.1..3.0) multiplies misfit evaluations, not syntheses:
    minimizer_engine.f90:511-521 `only_moment_changed`.  Same numbers as synthesising every candidate."""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources(stype, base)
    sc.set_refs_from(o, [g, o], ncomps)
    for e in (g, o):
        e.set_misfit_method(norm)
        if norm.startswith("floating"):
            e.set_floating_shiftrange(-0.3, 0.3)
        if stype == "eikonal":      # untapered norms of a folded synthetic inherit its noise-dependent strip end (see the fold test)
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
    p = np.tile(base, (8, 1))
    p[[1, 4, 6], 5] += 20                       # a second geometry, interleaved with the first
    p[:, 4] *= np.array([1.0, 0.6, 1.4, 2.0, 1.0, 0.8, 1.7, 0.3], np.float32)
    mg, sg = g.eval_sources(stype, p)
    t_shared = g.last_timing()
    mo, so = o.eval_sources(stype, p)
    assert not sg.any() and not so.any()
    tol = misfit_tol(mo, 0.25 if norm.startswith("ampspec") else 0.1)
    assert np.all(np.abs(mg - mo) <= tol), np.abs((mg - mo) / tol).max()
    g.set_share_syntheses(False)
    md, sd = g.eval_sources(stype, p)
    assert np.array_equal(mg, md) and np.array_equal(sg, sd)      # identical arithmetic per candidate, shared or not
    assert t_shared["launches"][2] >= 1


# circular: time north east depth moment strike dip rake radius rupture-velocity rise-time (source_circular.f90)
CIRC = np.array([0.2, 100, -150, 3000, 1.2e18, 60, 70, 30, 1200, 2800, 0.4], np.float32)
# point_lp: time north east depth moment mxx myy mzz mxy mxz myz duration-of-excitation period (source_point_lp.f90)
PLP = np.array([0.1, 50, -80, 2500, 1.0, 1e17, -0.4e17, -0.6e17, 0.3e17, 0.2e17, -0.5e17, 9.0, 2.0], np.float32)


@pytest.mark.parametrize("stype,params", [("circular", CIRC), ("point_lp", PLP)])
@pytest.mark.parametrize("norm", ["l2norm", "ampspec_l1norm"])
def test_circular_and_point_lp_sources(stype, params, norm):
    """the remaining two source types of source_all.f90:216-261: centroid table bit-exact
    (point_lp has 46 time centroids: more than one device group), synthetics and misfits within the bars"""
    g, o = engines(sc.small_db(), COMPS6)
    tg, gg, ng = g.discretize_source(stype, params)
    to, go, no = o.discretize_source(stype, params)
    assert ng == no and ng > 40 and np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    o.set_source_params(stype, params)
    g.set_source_params(stype, params)
    for ir in range(1, 7):
        for ic in range(1, len(COMPS6[ir - 1]) + 1):
            assert_seis_close(g.get_seismogram(ir, ic, 1), o.get_seismogram(ir, ic, 1), "rcv %d comp %d" % (ir, ic))
    sc.set_refs_from(o, [g, o], [len(c) for c in COMPS6])
    for e in (g, o):
        e.set_misfit_method(norm)
        for ir in range(1, 7):
            e.set_misfit_taper(ir, *TAPER)
    p = np.tile(params, (4, 1))
    p[1, 1] += 300; p[2, 3] -= 400; p[3, 4] *= 1.5
    if stype == "circular":
        p[2, 8] = 1500
    mg, sg = g.eval_sources(stype, p)
    mo, so = o.eval_sources(stype, p)
    assert not sg.any() and not so.any()
    tol = misfit_tol(mo, 0.25 if norm.startswith("ampspec") else 0.1)
    assert np.all(np.abs(mg - mo) <= tol), np.abs((mg - mo) / tol).max()


def test_minimize_lm_against_the_sequential_restatement():
    """minimize_lm (minimizer_engine.f90:729-874): lmdif on the per-trace misfits of the masked, normalised parameters, the
    Jacobian's finite-difference sources evaluated as one GPU batch.  Same start, same settings: both arrive at the source
    the references were made from.  The two trajectories are not bitwise the same (the misfits agree to 1e-5, and
    Levenberg-Marquardt amplifies that), so the comparison is on the end point, with the evaluation counts reported."""
    g, o = engines(sc.small_db(), COMPS6)
    truth = np.array(sc.BILAT_SMALL[0] if np.ndim(sc.BILAT_SMALL) > 1 else sc.BILAT_SMALL, np.float32)
    o.eval_sources("bilateral", truth)
    sc.set_refs_from(o, [g, o], [len(c) for c in COMPS6], scale=1.0)
    start = truth.copy()
    start[0] += 0.3; start[1] += 400.0; start[2] -= 300.0; start[3] += 250.0          # time north east depth
    mask = np.zeros(14, bool); mask[:4] = True
    res = {}
    for name, e in (("gpu", g), ("oracle", o)):
        e.set_misfit_method("l2norm")
        for ir in range(1, 7):
            e.set_misfit_taper(ir, *TAPER)
        e.set_source_params("bilateral", start)
        e.set_source_params_mask(mask)
        assert np.array_equal(e.get_source_subparams(), start[:4])
        e.set_source_subparams_limits(truth[:4] - np.array([2, 3000, 3000, 1500], np.float32), truth[:4] + np.array([2, 3000, 3000, 1500], np.float32))
        info, iterations, misfit = e.minimize_lm()
        res[name] = (info, iterations, misfit, e.get_source_subparams())
    (ig, ng, mg, xg), (io, no, mo, xo) = res["gpu"], res["oracle"]
    assert 1 <= ig <= 7 and 1 <= io <= 7, (ig, io)
    assert mg < 2e-3 and mo < 2e-3, (mg, mo)                       # started at a misfit of order 1
    scale = np.array([1.0, 10000.0, 10000.0, 10000.0])
    assert np.all(np.abs(xg - truth[:4]) / scale < 2e-3) and np.all(np.abs(xo - truth[:4]) / scale < 2e-3), (xg, xo, truth[:4])
    assert ng <= 2 * no + 10 and no <= 2 * ng + 10, (ng, no)
    # the engine is left at the last model evaluated and answers for it
    assert abs(g.get_global_misfit() - mg) <= 1e-6 * max(mg, 1e-3)


@pytest.mark.parametrize("taper", [False, True])
@pytest.mark.parametrize("stype,params", [("bilateral", sc.BILAT_SMALL), ("moment_tensor", sc.MT_SMALL)])
def test_peak_amplitudes_and_arias_intensities(stype, params, taper):
    """get_peak_amplitudes / get_arias_intensities (minimizer_engine.f90:1174-1245): peak vector norm of velocity and acceleration
    and Arias intensity of the synthetics per receiver, over the components receiver.f90:505-542 picks (three, the horizontal
    pair, the vertical alone: the six receivers of COMPS6 cover all cases); single source and batched."""
    g, o = engines(sc.small_db(), COMPS6)
    if taper:
        for e in (g, o):
            for ir in range(1, 7):
                e.set_misfit_taper(ir, *TAPER)
    g.set_source_params(stype, params)
    want = [o.get_ground_motion(stype, params, w) for w in (1, 2, 3)]
    got = [g.get_peak_amplitudes(1), g.get_peak_amplitudes(2), g.get_arias_intensities()]
    # first differences of traces that agree to ~1e-6 of their peak: 2e-5; second differences (acceleration, and its sum of
    # squares in the Arias intensity) take the same absolute trace error relative to a much smaller quantity: 2e-4
    rtol = {1: 2e-5, 2: 2e-4, 3: 2e-4}
    for w, a, b in zip((1, 2, 3), got, want):
        assert a.shape == b.shape == (6,) and np.all(b > 0)
        assert np.all(np.abs(a - b) <= rtol[w] * np.abs(b)), (w, a, b)
    p = np.tile(params, (3, 1)); p[1, 1] += 250; p[2, 3] += 300
    vals, st = g.eval_ground_motion(stype, p, 6)
    assert not st.any() and np.array_equal(vals[0], np.stack(got, 1))
    for i in (1, 2):
        for w in (1, 2, 3):
            b = o.get_ground_motion(stype, p[i], w)
            assert np.all(np.abs(vals[i, :, w - 1] - b) <= rtol[w] * np.abs(b)), (i, w)
    with pytest.raises(Exception, match="differentiate argument must be 1"):
        g.get_peak_amplitudes(3)


def test_status_does_not_depend_on_batch_composition_or_fast_path():
    """per-candidate failure statuses with the shared-synthesis and tensor-core fast paths on and off: a candidate with a
    non-finite moment listed first must not take the candidates that share its geometry with it, and a candidate whose
    misfits come out NaN reports status 2 on every path"""
    g, o = engines(sc.small_db(), COMPS6)
    ncomps = [len(c) for c in COMPS6]
    o.eval_sources("circular", CIRC)
    sc.set_refs_from(o, [g, o], ncomps)
    p = np.tile(CIRC, (4, 1))
    p[0, 4] = np.nan; p[2, 4] *= 1.3; p[3, 5] += 10          # NaN moment first; same geometry, other moment; other strike
    g.set_share_syntheses(True)
    ms, ss = g.eval_sources("circular", p)
    g.set_share_syntheses(False)
    md, sd = g.eval_sources("circular", p)
    mo, so = o.eval_sources("circular", p)
    assert list(ss) == list(sd) == [1, 0, 0, 0], (ss, sd, so)      # (the product refuses to discretise with a non-finite parameter)
    assert not so[1:].any()
    assert np.all(np.abs(ms[1:] - mo[1:]) <= misfit_tol(mo[1:])) and np.all(np.abs(md[1:] - mo[1:]) <= misfit_tol(mo[1:]))
    # bilateral sources take any moment: the NaN comes out of the misfits (status 2), again for that candidate only
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps)
    q = np.tile(sc.BILAT_SMALL, (3, 1)); q[0, 4] = np.nan; q[2, 4] *= 0.5
    g.set_share_syntheses(True)
    _, s1 = g.eval_sources("bilateral", q)
    g.set_share_syntheses(False)
    _, s2 = g.eval_sources("bilateral", q)
    assert list(s1) == list(s2) == [2, 0, 0]
    # moment-tensor grid: one tensor with a NaN component
    o.eval_sources("moment_tensor", sc.MT_SMALL)
    sc.set_refs_from(o, [g, o], ncomps)
    r = _mt_grid()
    r[5, 6] = np.nan
    g.set_mt_grid(True)
    mg, sg = g.eval_sources("moment_tensor", r)
    assert g.last_timing()["launches"][3] >= 1
    g.set_mt_grid(False)
    md, sd = g.eval_sources("moment_tensor", r)
    assert np.array_equal(sg, sd) and sg[5] == 2 and sg.sum() == 2
