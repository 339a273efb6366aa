"""Diagnostic (not a test): one seed of the random sweep in detail -- misfits of GPU / fp32 oracle / wide oracle and the seismograms of
the worst candidate.   PYTHONPATH=.:tests python scratch/diag_random_seed.py 1064 404 ..."""
import sys
import numpy as np
import scenario as sc
from oracle_lib import OracleEngine
from test_random_parity_gpu import random_case, RTOL
from kiwi_b200 import Engine

for seed in [int(a) for a in sys.argv[1:]]:
    lat, lon, dep, comps, stype, base, cands, cfg = random_case(seed)
    print("=== seed", seed, stype, cfg, "comps", comps, "dep", dep)
    db = getattr(sc, cfg["db"])()
    g, o, w = Engine(0), OracleEngine(), OracleEngine(wide=True)
    for e in (g, o, w):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    o.set_source_params(stype, base)
    ncomps = [len(c) for c in comps]
    refs = sc.set_refs_from(o, [g, o], ncomps)
    for (ir, ic), (first, data) in refs.items():
        w.set_ref_seismogram(ir, ic, (first - 1) * 0.1, data)
    for e in (g, o, w):
        e.set_misfit_method(cfg["norm"]); e.set_synthetics_factor(cfg["factor"])
        if cfg["norm"].startswith("floating"):
            e.set_floating_shiftrange(-0.4, 0.3)
        if cfg["taper"]:
            for ir in range(1, len(comps) + 1):
                e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
        if cfg["filt"]:
            e.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
        if cfg["disable"]:
            e.switch_receiver(cfg["disable"], False)
    if cfg.get("autoshift"):
        print("  (autoshift case: shifts not applied in this diagnostic)")
    mg, sg = g.eval_sources(stype, cands)
    mo, so = o.eval_sources(stype, cands)
    mw, sw = w.eval_sources(stype, cands)
    floor = 0.25 if cfg["norm"].startswith("ampspec") else 0.1
    tol = RTOL * np.maximum(np.abs(mo), floor * np.abs(mo[..., 1:2]))
    lim = np.maximum(tol, 2.0 * np.abs(mo - mw))
    ratio = np.abs(mg - mo) / lim
    ic, im, ik = np.unravel_index(np.argmax(ratio), ratio.shape)
    print("  worst ratio %.3f at candidate %d misfit slot %d (%s): gpu %.9g fp32 %.9g wide %.9g  norm factor %.6g  |g-o|/nf %.3g |g-w|/nf %.3g |o-w|/nf %.3g"
          % (ratio.max(), ic, im, ["misfit", "norm"][ik], mg[ic, im, ik], mo[ic, im, ik], mw[ic, im, ik], mo[ic, im, 1],
             abs(mg[ic, im, ik] - mo[ic, im, ik]) / mo[ic, im, 1], abs(mg[ic, im, ik] - mw[ic, im, ik]) / mo[ic, im, 1], abs(mo[ic, im, ik] - mw[ic, im, ik]) / mo[ic, im, 1]))
    print("  candidate", ic, "params", cands[ic], "status", sg, so)
    g.set_source_params(stype, cands[ic])
    for e in (o, w):
        e.eval_sources(stype, cands[ic:ic + 1])
    for ir in range(1, len(comps) + 1):
        if ir == cfg["disable"]:
            continue
        for k in range(1, ncomps[ir - 1] + 1):
            (fg, dg), (fo, do), (fw, dw) = g.get_seismogram(ir, k), o.get_seismogram(ir, k), w.get_seismogram(ir, k)
            if fg != fo or dg.size != do.size:
                print("  rcv %d comp %d: SPAN differs gpu [%d,+%d) oracle [%d,+%d)" % (ir, k, fg, dg.size, fo, do.size)); continue
            pk = np.abs(do).max()
            j = int(np.argmax(np.abs(dg - do)))
            print("  rcv %d comp %d: span [%d,+%d)  |g-o|/peak %.3g (at sample %d of %d)  |g-w| %.3g  |o-w| %.3g   peak %.4g  rms %.4g"
                  % (ir, k, fg, dg.size, np.abs(dg - do).max() / pk, j, dg.size, np.abs(dg - dw).max() / pk, np.abs(do - dw).max() / pk, pk, np.sqrt((do ** 2).sum())))
            for proc in ("plain", "tapered", "filtered"):
                try:
                    pg, po = g.get_probe(ir, k, "synthetics", proc), o.get_probe(ir, k, "synthetics", proc)
                    if pg[1].size == po[1].size:
                        print("      probe %-8s first %d/%d n %d  |g-o|/peak %.3g  l2 of diff / l2 %.3g" % (proc, pg[0], po[0], pg[1].size, np.abs(pg[1] - po[1]).max() / np.abs(po[1]).max(),
                              np.sqrt(((pg[1] - po[1]) ** 2).sum() / (po[1] ** 2).sum())))
                    else:
                        print("      probe %-8s sizes differ %d %d" % (proc, pg[1].size, po[1].size))
                except Exception as exc:
                    print("      probe %-8s: %s" % (proc, exc))
