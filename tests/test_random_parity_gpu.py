"""Seeded random sweep: random sources, receiver geometries (incl. receiver depth, components, disabled receivers,
centroids that leave the database), interpolation / undersampling settings, norms, tapers, filters and factors --
CUDA path against the oracle on every case."""
import os

import numpy as np
import pytest

import scenario as sc
from oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu
RTOL = 1e-5
ALL_COMPS = ["ned", "swu", "ar", "cl", "d", "u", "ne", "rd", "nedar", "wsucl", "a", "e"]
NORMS = ["l2norm", "l1norm", "scalar_product", "peak", "ampspec_l2norm", "ampspec_l1norm", "floating_l2norm", "floating_l1norm"]


def random_case(seed):
    rng = np.random.default_rng(seed)
    nr = int(rng.integers(2, 7))
    lat, lon, dep = sc.small_receivers(nr, seed=int(rng.integers(1, 10 ** 6)), dmin=float(rng.uniform(1.5e3, 9e3)), dmax=float(rng.uniform(10e3, 17e3)))
    dep = rng.choice([0.0, 0.0, 250.0, 600.0], nr).astype(np.float32)
    comps = [ALL_COMPS[int(i)] for i in rng.integers(0, len(ALL_COMPS), nr)]
    stype = ["bilateral", "moment_tensor", "eikonal", "bilateral"][int(rng.integers(0, 4))]
    if stype == "bilateral":
        base = np.array([rng.uniform(-0.5, 1.0), rng.uniform(-800, 800), rng.uniform(-800, 800), rng.uniform(1200, 4200), 10 ** rng.uniform(17, 19),
                         rng.uniform(0, 360), rng.uniform(5, 90), rng.uniform(-180, 180), rng.uniform(-90, 90), rng.uniform(0, 3000),
                         rng.uniform(0, 3000), rng.uniform(0, 2500), rng.uniform(2000, 3500), rng.uniform(0, 1.0)], np.float32)
        if rng.random() < 0.2:
            base[9:12] = 0          # point-like: nx = ny = 1
    elif stype == "moment_tensor":
        base = np.concatenate([[rng.uniform(-0.5, 1.0), rng.uniform(-800, 800), rng.uniform(-800, 800), rng.uniform(800, 4800)],
                               rng.normal(0, 1e18, 6), [rng.uniform(0.05, 1.2)]]).astype(np.float32)
    else:
        base = np.array([rng.uniform(-0.3, 0.5), rng.uniform(-500, 500), rng.uniform(-500, 500), rng.uniform(2500, 4000), 10 ** rng.uniform(17, 19),
                         rng.uniform(0, 360), rng.uniform(20, 90), rng.uniform(-180, 180), rng.uniform(-300, 300), rng.uniform(-300, 300),
                         rng.uniform(600, 2000), 0, 0, rng.uniform(0.6, 0.95), rng.choice([0.0, 0.3, 0.6])], np.float32)
    cfg = dict(interp=["bilinear", "nearest_neighbor"][int(rng.random() < 0.25)], under=(int(rng.integers(1, 3)), int(rng.integers(1, 3))),
               eff_dt=float(rng.choice([0.1, 0.2, 0.35])), norm=NORMS[int(rng.integers(0, len(NORMS)))], taper=bool(rng.random() < 0.6),
               filt=bool(rng.random() < 0.4), factor=float(rng.choice([1.0, 1.0, 0.8])), disable=int(rng.integers(0, nr + 1)),
               db=["small_db", "small_db_ng8"][int(rng.random() < 0.2)])
    if seed >= 24:   # later seeds also draw the remaining source types (separate generator: seeds 0..23 stay what they were)
        rng2 = np.random.default_rng(10 ** 6 + seed)
        u = rng2.random()
        if u < 0.15:
            stype = "circular"
            base = np.array([rng2.uniform(-0.3, 0.6), rng2.uniform(-500, 500), rng2.uniform(-500, 500), rng2.uniform(2200, 4000), 10 ** rng2.uniform(17, 19),
                             rng2.uniform(0, 360), rng2.uniform(10, 90), rng2.uniform(-180, 180), rng2.uniform(300, 1500), rng2.uniform(2000, 3200),
                             rng2.uniform(0, 0.8)], np.float32)
        elif u < 0.3:
            stype = "point_lp"
            base = np.concatenate([[rng2.uniform(-0.3, 0.6), rng2.uniform(-500, 500), rng2.uniform(-500, 500), rng2.uniform(1500, 4000), 1.0],
                                   rng2.normal(0, 1e17, 6), [rng2.uniform(3.0, 10.0), rng2.uniform(1.0, 3.0)]]).astype(np.float32)
        cfg["autoshift"] = bool(rng2.random() < 0.3)
    if stype == "eikonal":
        cfg["taper"] = True          # untapered norms of folded synthetics: see test_eikonal_seismograms_with_rise_time_fold
    n = int(rng.integers(2, 6))
    cands = np.tile(base, (n, 1))
    for i in range(1, n):
        k = int(rng.integers(0, 6)) if stype not in ("moment_tensor", "point_lp") else int(rng.integers(1, 10))
        if stype == "point_lp" and k == 4:
            k = 5
        cands[i, k] += np.float32(rng.normal(0, 1) * (10.0 if k >= 5 and stype not in ("moment_tensor", "point_lp") else (200.0 if k in (1, 2, 3) else (0.2 if k == 0 else abs(base[k]) * 0.3))))
    return lat, lon, dep, comps, stype, base, cands, cfg


@pytest.mark.parametrize("seed", range(int(os.environ.get("KIWI_RANDOM_CASES", "48"))))
def test_random_case(seed):
    from kiwi_b200 import Engine
    lat, lon, dep, comps, stype, base, cands, cfg = random_case(seed)
    db = getattr(sc, cfg["db"])()
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    to, go, no = o.discretize_source(stype, base)
    tg, gg, ng = g.discretize_source(stype, base)
    assert ng == no and np.array_equal(tg.view(np.uint32), to.view(np.uint32)), "centroid table"
    o.set_source_params(stype, base)
    ncomps = [len(c) for c in comps]
    try:
        refs = sc.set_refs_from(o, [g, o], ncomps)
    except Exception:
        pytest.skip("base source leaves the database at some receiver (no synthetic to use as reference)")

    def configure(e):
        e.set_misfit_method(cfg["norm"])
        e.set_synthetics_factor(cfg["factor"])
        if cfg["norm"].startswith("floating"):
            e.set_floating_shiftrange(-0.4, 0.3)
        if cfg["taper"]:
            for ir in range(1, len(comps) + 1):
                e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
        if cfg["filt"]:
            e.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
        if cfg["disable"]:
            e.switch_receiver(cfg["disable"], False)

    for e in (g, o):
        configure(e)
    if g.nmisfits == 0:
        pytest.skip("all receivers disabled")
    applied = np.zeros(len(comps), np.float32)
    if cfg.get("autoshift"):   # autoshift_ref_seismogram on the base source: the same integer shifts, unless the correlation peak is a near tie
        g.set_source_params(stype, base)
        en = [ir for ir in range(1, len(comps) + 1) if ir != cfg["disable"]]
        cc = [(g.get_cross_correlations(ir, -0.3, 0.4), o.get_cross_correlations(ir, -0.3, 0.4)) for ir in en]
        for cg, co in cc:
            assert np.all(np.abs(cg - co) <= 4 * RTOL * max(np.abs(co).max(), 1e-30))
        sg_, so_ = g.autoshift_ref_seismogram(0, -0.3, 0.4), o.autoshift_ref_seismogram(0, -0.3, 0.4)
        for k, ir in enumerate(en):
            co = cc[k][1]
            score = (np.maximum(co / max(1.0, co.max()), 0.0) ** 2).sum(0)
            top = np.sort(score)[::-1]
            if top.size > 1 and top[0] - top[1] > 1e-3 * top[0]:
                assert sg_[ir - 1] == so_[ir - 1], (ir, sg_, so_)
            else:
                g.shift_ref_seismogram(ir, float(so_[ir - 1] - sg_[ir - 1]))    # near tie: continue from the oracle's choice
        applied = so_
    mg, sg = g.eval_sources(stype, cands)
    mo, so = o.eval_sources(stype, cands)
    assert np.array_equal(sg > 0, so > 0), (sg, so)
    ok = so == 0
    floor = 0.25 if cfg["norm"].startswith("ampspec") else 0.1

    def tolerance(m, floor=floor):
        if cfg["norm"] in ("scalar_product", "peak"):
            return RTOL * np.maximum(np.abs(m), floor * np.abs(m).max(axis=(0, 1), keepdims=True))
        return RTOL * np.maximum(np.abs(m), floor * np.abs(m[..., 1:2]))

    tol = tolerance(mo)
    if np.all(np.abs(mg[ok] - mo[ok]) <= tol[ok]):
        return
    # beyond the tight bar against the fp32 restatement (1e-5 of the misfit, or of 0.1 / 0.25 of the norm factor for small misfits): the
    # bar is then what 1e-5 agreement of the SEISMOGRAMS implies for a norm of their difference -- 1e-5 of the norm factor -- against the
    # restatement with strips carried in double, and twice the fp32 restatement's own distance from it against the fp32 one
    # (accumulation noise).  About 3 % of the seeded cases come here.
    w = OracleEngine(wide=True)
    sc.setup(w, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    for (ir, ic), (first, data) in refs.items():
        w.set_ref_seismogram(ir, ic, (first - 1) * 0.1, data)
    configure(w)
    for ir in range(1, len(comps) + 1):
        if applied[ir - 1] != 0.0:
            w.shift_ref_seismogram(ir, float(applied[ir - 1]))
    mw, sw = w.eval_sources(stype, cands)
    assert np.array_equal(sw > 0, so > 0)
    tolw = tolerance(mw, 1.0)
    assert np.all(np.abs(mg[ok] - mw[ok]) <= tolw[ok]), (cfg, stype, float(np.abs((mg[ok] - mw[ok]) / tolw[ok]).max()))
    lim = np.maximum(tol, 2.0 * np.abs(mo - mw))
    if np.all(np.abs(mg[ok] - mo[ok]) <= lim[ok]):
        return
    # Still beyond (8 of 1500 seeds, by up to 3.7 x, profiles/r02_random_sweep.md): components that are small sums of large terms --
    # a transverse or near-nodal trace 20-50 x smaller than the tensor terms it is made of -- carry the rounding noise of those terms,
    # ~3e-6 of their own peak in either implementation, and one draw of that noise (the fp32 restatement's) is no bound for another
    # (the batched order of operations).  There the bar is the one seismogram parity implies, 1e-5 of the norm factor, for the batched
    # path, and the tight bar for the reference's order of operations (kiwi_set_accumulation), whose seismograms are the restatement's.
    tol1 = tolerance(mo, 1.0)
    assert np.all(np.abs(mg[ok] - mo[ok]) <= tol1[ok]), (cfg, stype, float(np.abs((mg[ok] - mo[ok]) / tol1[ok]).max()))
    g.set_accumulation(True)
    mr, sr = g.eval_sources(stype, cands)
    assert np.array_equal(sr > 0, so > 0)
    assert np.all(np.abs(mr[ok] - mo[ok]) <= tol[ok]), ("reference order", cfg, stype, float(np.abs((mr[ok] - mo[ok]) / tol[ok]).max()))


def random_grid_case(seed):
    rng = np.random.default_rng(5000 + seed)
    nr = int(rng.integers(2, 7))
    lat, lon, dep = sc.small_receivers(nr, seed=int(rng.integers(1, 10 ** 6)), dmin=float(rng.uniform(3e3, 9e3)), dmax=float(rng.uniform(10e3, 16e3)))
    comps = [ALL_COMPS[int(i)] for i in rng.integers(0, len(ALL_COMPS), nr)]
    nloc, nmt = int(rng.integers(1, 6)), int(rng.integers(8, 160))
    mts = rng.normal(0, 1e18, (nmt, 6)).astype(np.float32)
    if rng.random() < 0.3:
        mts[int(rng.integers(0, nmt))] = 0.0                 # a zero tensor among the candidates
    p = np.zeros((nloc, nmt, 11), np.float32)
    for l in range(nloc):
        p[l, :, 0] = rng.uniform(-0.3, 0.8); p[l, :, 1] = rng.uniform(-700, 700); p[l, :, 2] = rng.uniform(-700, 700)
        p[l, :, 3] = rng.uniform(900, 4200); p[l, :, 4:10] = mts; p[l, :, 10] = rng.choice([0.0, 0.3, 0.7, 1.1])
    p = p.reshape(-1, 11)
    p = p[rng.permutation(p.shape[0])]
    if rng.random() < 0.5:
        p = p[: max(8, int(p.shape[0] * rng.uniform(0.5, 1.0)))]      # ragged: locations with different numbers of tensors
    cfg = dict(norm=["l2norm", "l1norm"][int(rng.integers(0, 2))], taper=bool(rng.random() < 0.5), factor=float(rng.choice([1.0, 1.0, 0.7])),
               eff_dt=float(rng.choice([0.2, 0.5])), disable=int(rng.integers(0, nr + 1)), interp=["bilinear", "nearest_neighbor"][int(rng.random() < 0.2)],
               db=["small_db", "small_db_ng8"][int(rng.random() < 0.25)])
    return lat, lon, dep, comps, p, cfg


@pytest.mark.parametrize("seed", range(int(os.environ.get("KIWI_RANDOM_GRID_CASES", "16"))))
def test_random_moment_tensor_grid(seed):
    """point moment-tensor grid searches of random shape through the tcgen05 contraction: against the oracle and the direct path"""
    from kiwi_b200 import Engine
    lat, lon, dep, comps, p, cfg = random_grid_case(seed)
    db = getattr(sc, cfg["db"])()
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"])
    o.set_source_params("moment_tensor", sc.MT_SMALL)
    try:
        sc.set_refs_from(o, [g, o], [len(c) for c in comps])
    except Exception:
        pytest.skip("the reference source leaves the database")
    for e in (g, o):
        e.set_misfit_method(cfg["norm"]); e.set_synthetics_factor(cfg["factor"])
        if cfg["taper"]:
            for ir in range(1, len(comps) + 1):
                e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
        if cfg["disable"]:
            e.switch_receiver(cfg["disable"], False)
    if g.nmisfits == 0:
        pytest.skip("all receivers disabled")
    mg, sg = g.eval_sources("moment_tensor", p)
    used_grid = g.last_timing()["launches"][3] >= 1
    g.set_mt_grid(2)                      # synthesis not fused into the contraction (k_synth + k_mt_contract)
    mu, su = g.eval_sources("moment_tensor", p)
    g.set_mt_grid(False)
    md, sd = g.eval_sources("moment_tensor", p)
    mo, so = o.eval_sources("moment_tensor", p)
    assert np.array_equal(sg > 0, so > 0) and np.array_equal(sd > 0, so > 0) and np.array_equal(su > 0, so > 0), (sg, su, sd, so)
    ok = so == 0
    tol = RTOL * np.maximum(np.abs(mo), 0.1 * np.abs(mo[..., 1:2]))
    assert np.all(np.abs(md[ok] - mo[ok]) <= tol[ok]), ("direct", cfg, float(np.abs((md[ok] - mo[ok]) / tol[ok]).max()))
    assert np.all(np.abs(mg[ok] - mo[ok]) <= tol[ok]), ("grid" if used_grid else "direct (no grid found)", cfg, float(np.abs((mg[ok] - mo[ok]) / tol[ok]).max()))
    assert np.all(np.abs(mu[ok] - mo[ok]) <= tol[ok]), ("unfused grid" if used_grid else "direct (no grid found)", cfg, float(np.abs((mu[ok] - mo[ok]) / tol[ok]).max()))


@pytest.mark.parametrize("seed", range(int(os.environ.get("KIWI_RANDOM_SHARE_CASES", "8"))))
def test_random_shared_syntheses(seed):
    """batches in which random subsets of candidates differ only in the moment (the batched only_moment_changed shortcut):
    bit-identical to evaluating every candidate on its own"""
    from kiwi_b200 import Engine
    rng = np.random.default_rng(9000 + seed)
    lat, lon, dep, comps, stype, base, cands, cfg = random_case(int(rng.integers(24, 400)))
    if stype == "moment_tensor":
        stype, base = "bilateral", sc.BILAT_SMALL.copy()
    db = getattr(sc, cfg["db"])()
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    o.set_source_params(stype, base)
    try:
        sc.set_refs_from(o, [g], [len(c) for c in comps])
    except Exception:
        pytest.skip("base source leaves the database")
    g.set_misfit_method(cfg["norm"] if not cfg["norm"].startswith("floating") else "l2norm")
    n = int(rng.integers(3, 12))
    distinct = np.tile(base, (int(rng.integers(1, 4)), 1))
    for i in range(1, distinct.shape[0]):
        distinct[i, 5 if stype != "point_lp" else 1] += np.float32(7.0 * i)
    p = distinct[rng.integers(0, distinct.shape[0], n)].copy()
    p[:, 4] *= rng.uniform(0.2, 3.0, n).astype(np.float32)
    shared, s1 = g.eval_sources(stype, p)
    g.set_share_syntheses(False)
    single, s2 = g.eval_sources(stype, p)
    assert np.array_equal(s1, s2) and np.array_equal(shared.view(np.uint32), single.view(np.uint32))


@pytest.mark.parametrize("seed", range(int(os.environ.get("KIWI_RANDOM_ORDER_CASES", "24"))))
def test_random_case_in_the_references_order_of_operations(seed):
    """the random cases through the reference-order synthesis (kiwi_set_accumulation): the seismograms of the base source are bit-identical
    to the oracle's for every receiver and component, whatever the source type, components, receiver depth, interpolation / undersampling
    and database (8 or 10 components)"""
    from kiwi_b200 import Engine
    lat, lon, dep, comps, stype, base, cands, cfg = random_case(7000 + seed)
    base = base.copy()
    if stype == "eikonal":
        base[14] = 0.0      # no rise-time fold: the end of a folded strip is decided by fp32 noise in the reference
    if stype == "circular":
        base[10] = 0.0
    db = getattr(sc, cfg["db"])()
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    g.set_accumulation(True)
    try:
        o.eval_sources(stype, base)
    except Exception:
        pytest.skip("the oracle refuses the source")
    try:
        g.set_source_params(stype, base)
        g.get_seismogram(1, 1)
    except Exception:
        pytest.skip("the source cannot be evaluated (it leaves the database or its parameters are refused)")
    nsame = ntot = 0
    for ir in range(1, len(comps) + 1):
        for ic in range(1, len(comps[ir - 1]) + 1):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic)
            assert (fg, dg.size) == (fo, do.size), (stype, cfg, ir, ic, fg, dg.size, fo, do.size)
            nsame += int((dg.view(np.uint32) == do.view(np.uint32)).sum()); ntot += dg.size
    assert ntot > 0 and nsame == ntot, (stype, cfg, nsame, ntot)
