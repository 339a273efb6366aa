"""one-off diagnostic of the extended random sweep: per failing seed, where and by how much"""
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import scenario as sc
import test_random_parity_gpu as t
from oracle_lib import OracleEngine
from kiwi_b200 import Engine
RTOL = 1e-5
seeds = [int(x) for x in sys.argv[1:]] or list(range(24, 400))
for seed in seeds:
    lat, lon, dep, comps, stype, base, cands, cfg = t.random_case(seed)
    db = getattr(sc, cfg["db"])()
    g, o = Engine(0), OracleEngine()
    for e in (g, o):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    o.set_source_params(stype, base)
    ncomps = [len(c) for c in comps]
    try:
        sc.set_refs_from(o, [g, o], ncomps)
    except Exception:
        continue
    for e in (g, o):
        e.set_misfit_method(cfg["norm"]); e.set_synthetics_factor(cfg["factor"])
        if cfg["norm"].startswith("floating"): e.set_floating_shiftrange(-0.4, 0.3)
        if cfg["taper"]:
            for ir in range(1, len(comps) + 1): e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
        if cfg["filt"]: e.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
        if cfg["disable"]: e.switch_receiver(cfg["disable"], False)
    if g.nmisfits == 0: continue
    tag = "seed %d %s %s taper=%d filt=%d fac=%g interp=%s under=%s dis=%d db=%s" % (seed, stype, cfg["norm"], cfg["taper"], cfg["filt"], cfg["factor"], cfg["interp"], cfg["under"], cfg["disable"], cfg["db"])
    if cfg.get("autoshift"):
        g.set_source_params(stype, base)
        en = [ir for ir in range(1, len(comps) + 1) if ir != cfg["disable"]]
        for ir in en:
            cg, co = g.get_cross_correlations(ir, -0.3, 0.4), o.get_cross_correlations(ir, -0.3, 0.4)
            err = np.abs(cg - co).max() / max(np.abs(co).max(), 1e-30)
            if err > 4 * RTOL:
                print("XCORR", tag, "rcv", ir, comps[ir - 1], "err %.3g" % err, "cg", cg[0, :4], "co", co[0, :4], flush=True)
        sg_, so_ = g.autoshift_ref_seismogram(0, -0.3, 0.4), o.autoshift_ref_seismogram(0, -0.3, 0.4)
        if not np.array_equal(sg_, so_):
            print("SHIFT", tag, sg_, so_, flush=True)
            for ir in en: g.shift_ref_seismogram(ir, float(so_[ir - 1] - sg_[ir - 1]))
    mg, sg = g.eval_sources(stype, cands)
    mo, so = o.eval_sources(stype, cands)
    if not np.array_equal(sg > 0, so > 0):
        print("STATUS", tag, sg, so, flush=True); continue
    ok = so == 0
    floor = 0.25 if cfg["norm"].startswith("ampspec") else 0.1
    if cfg["norm"] in ("scalar_product", "peak"):
        tol = RTOL * np.maximum(np.abs(mo), floor * np.abs(mo).max(axis=(0, 1), keepdims=True))
    else:
        tol = RTOL * np.maximum(np.abs(mo), floor * np.abs(mo[..., 1:2]))
    r = np.abs(mg[ok] - mo[ok]) / tol[ok]
    if r.size and r.max() > 1:
        i = np.unravel_index(np.argmax(np.abs(mg - mo) / tol * ok[:, None, None]), mg.shape)
        print("MISFIT", tag, "ratio %.3g at %s gpu %.6g oracle %.6g nf %.6g" % (r.max(), i, mg[i], mo[i], mo[i[0], i[1], 1]), "autoshift" if cfg.get("autoshift") else "", flush=True)
print("done")
