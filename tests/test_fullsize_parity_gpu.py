"""Full-size parity (BASELINE.json configs C3 and C5 at their real sizes) of the CUDA path, through the C ABI,
against the oracle: GF indices / sample shifts / strip spans bit-exact (gfdb.f90:781-815, sparse_trace.f90:639-653)
outside the pairs the device flags as sitting on a cell edge, whose number is bounded; seismograms
(seismogram.f90:131-289) for every receiver and component

* within 1e-5 of the trace peak of the restatement with the strips carried in double (oracle -DKO_WIDE), and
* within 1e-5 + 1.5 x (fp32 restatement vs double-accumulated restatement) of the fp32 restatement: the reference adds
  ~1e4 sub-sources x 10 components one after the other into fp32 strips, which at this size leaves it 5e-5 of the peak
  away from the exact sum of the same terms (measured below); an implementation that sums in any other order cannot be
  closer to it than that.  Where the reference's own noise is below the bar (the same fault at effective_dt 1.0, ~1e3
  sub-sources) the fp32 restatement is matched to 1e-5 as it stands.

The misfit bar follows from the seismogram bar: with references at 1.07 x the synthetics a misfit is a 7 % difference
of the traces, so 1e-5 of the peak in the traces is ~1.5e-4 of such a misfit."""
import functools
import os

import numpy as np
import pytest

from oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # BASELINE.json north_star: "within 1e-5 relative for seismograms and misfits (fp32)"


def setup_pair(name, nrcv=None, wide=False, effective_dt=None):
    import bench
    from kiwi_b200 import Engine, synthetic
    w = dict(bench.WORKLOADS[name])
    if nrcv:
        w["nrcv"] = nrcv
    if effective_dt:
        w["effective_dt"] = effective_dt
    db = bench_l()
    lat, lon, dep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
    es = [Engine(0), OracleEngine(threads=os.cpu_count() or 1)] + ([OracleEngine(threads=os.cpu_count() or 1, wide=True)] if wide else [])
    for e in es:
        bench.configure(e, db, w, lat, lon, dep)
    return es + [w]


@functools.lru_cache(maxsize=None)
def bench_l():
    import bench
    from kiwi_b200 import synthetic
    w = bench.WORKLOADS["c3"]
    return synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"])


def seismogram_deviation(g, o, nrcv, ncomp=3, spans=True):
    """max over traces of max|g - o| / peak(o); spans must be equal"""
    worst = 0.0
    for ir in range(1, nrcv + 1):
        for ic in range(1, ncomp + 1):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic)
            assert (fg, dg.size) == (fo, do.size), "span of receiver %d component %d: [%d,+%d) vs [%d,+%d)" % (ir, ic, fg, dg.size, fo, do.size)
            peak = float(np.abs(do).max())
            assert peak > 0
            worst = max(worst, float(np.abs(dg - do).max()) / peak)
    return worst


def assert_seismograms(g, o, ow, nrcv, label):
    """the two-sided bar of the module docstring"""
    d_wide, d_f32, noise = seismogram_deviation(g, ow, nrcv), seismogram_deviation(g, o, nrcv), seismogram_deviation(o, ow, nrcv)
    print("%s: gpu vs double-accumulated %.2e, gpu vs fp32 path %.2e, fp32 path vs double-accumulated %.2e of the trace peak" % (label, d_wide, d_f32, noise))
    assert d_wide <= RTOL, "%s: seismograms deviate by %.3g of the trace peak from the double-accumulated restatement" % (label, d_wide)
    assert d_f32 <= RTOL + 1.5 * noise, "%s: %.3g from the fp32 restatement, whose own accumulation noise is %.3g" % (label, d_f32, noise)


def test_c3_seismograms_and_integers_match_the_fp32_oracle():
    """C3: the Izmit bilateral source at effective_dt 0.35 (10290 sub-sources) x 200 receivers x ned"""
    from kiwi_b200 import synthetic
    g, o, ow, w = setup_pair("c3", wide=True)
    o.record_indices(True)
    o.eval_sources("bilateral", synthetic.IZMIT)      # (no references: status 1, the seismograms are there)
    ow.eval_sources("bilateral", synthetic.IZMIT)
    g.set_source_params("bilateral", synthetic.IZMIT)
    table, grid, n = g.discretize_source("bilateral", synthetic.IZMIT)
    to, go, no = o.discretize_source("bilateral", synthetic.IZMIT)
    assert n == no == 10290 and list(grid) == list(go) == [98, 15, 7]
    assert np.array_equal(table.view(np.uint32), to.view(np.uint32))
    nflag = ndiff = npairs = 0
    for ir in range(1, w["nrcv"] + 1):
        ig, io = g.get_indices(ir), o.get_indices(ir)
        assert ig["ix"].size == io["ix"].size == 10290
        ok = ig["near"] == 0       # distance (fp64 libm: device vs glibc) within 4 fp32 ulps of a cell edge
        nflag += int((~ok).sum()); npairs += ok.size
        ndiff += int((ig["ix"] != io["ix"]).sum())
        assert np.array_equal(ig["ix"][ok], io["ix"][ok]), "ix of receiver %d" % ir
        assert np.array_equal(ig["iz"], io["iz"]) and np.array_equal(ig["its"], io["its"]), "iz / its of receiver %d" % ir
        assert np.array_equal(ig["diz"], io["diz"])
        # since the sub-source azimuths atan2f(east, north) come from the host library (round 2) the distance coordinate agrees to the last
        # bit as well, flagged pairs included: indices and bilinear weights are bit-exact everywhere
        assert np.array_equal(ig["ix"], io["ix"]) and np.array_equal(ig["dix"].view(np.uint32), io["dix"].view(np.uint32)), "ix / dix of receiver %d" % ir
    # a distance of ~1000 cells is within 4 ulps of an edge with probability ~1e-3
    assert nflag <= 4e-3 * npairs, "%d of %d (sub-source, receiver) pairs flagged as sitting on a cell edge" % (nflag, npairs)
    print("C3: %d of %d (sub-source, receiver) pairs flagged as sitting on a cell edge, %d distance indices differ" % (nflag, npairs, ndiff))
    assert ndiff == 0, "%d of %d GF distance indices differ" % (ndiff, npairs)
    assert_seismograms(g, o, ow, w["nrcv"], "C3")


def test_c3_fault_at_a_tenth_of_the_sub_sources_matches_the_fp32_oracle_as_it_stands():
    """the same fault and receivers at effective_dt 1.0 (~1e3 sub-sources): the fp32 path's accumulation noise is below the bar"""
    from kiwi_b200 import synthetic
    g, o, w = setup_pair("c3", effective_dt=1.0)
    o.eval_sources("bilateral", synthetic.IZMIT)
    g.set_source_params("bilateral", synthetic.IZMIT)
    table, grid, n = g.discretize_source("bilateral", synthetic.IZMIT)
    assert 500 < n < 3000
    worst = seismogram_deviation(g, o, w["nrcv"])
    assert worst <= RTOL, "%d sub-sources: seismograms deviate by %.3g of the trace peak" % (n, worst)


def test_c5_candidate_seismograms_match_the_fp32_oracle():
    """one candidate of the C5 sweep on the dense array: 2000 receivers x ned (every trace compared)"""
    from kiwi_b200 import synthetic
    g, o, ow, w = setup_pair("c5", wide=True)
    cand = synthetic.bilateral_sweep(32)[5]
    o.eval_sources("bilateral", cand)
    ow.eval_sources("bilateral", cand)
    g.set_source_params("bilateral", cand)
    assert_seismograms(g, o, ow, w["nrcv"], "C5 candidate")


def test_c3_misfits_follow_from_the_seismogram_bar():
    """misfits of a C3 batch against the fp32 oracle: within 1e-5 of the norm factor, i.e. of the size of the traces the
    misfit is a difference of (the reference is 1.07 x the base synthetics, so the misfits are ~0.07 x the norm factors)"""
    import bench
    from kiwi_b200 import synthetic
    g, o, ow, w = setup_pair("c3", wide=True)
    g.set_source_params("bilateral", synthetic.IZMIT)
    bench.set_references(g, [g, o, ow], w["nrcv"], bench_l().meta()["dt"])
    p = synthetic.bilateral_sweep(32)[:2]
    mg, sg = g.eval_sources("bilateral", p)
    mo, so = o.eval_sources("bilateral", p)
    mw, sw = ow.eval_sources("bilateral", p)
    assert not sg.any() and not so.any() and not sw.any()
    nf = np.abs(mw[..., 1:2])
    noise = float(np.max(np.abs(mo - mw) / nf))
    assert np.all(np.abs(mg - mw) <= RTOL * nf)
    assert np.all(np.abs(mg - mo) <= (RTOL + 1.5 * noise) * nf)
