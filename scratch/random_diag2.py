import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import scenario as sc
import test_random_parity_gpu as t
from oracle_lib import OracleEngine
from kiwi_b200 import Engine
np.set_printoptions(linewidth=220, precision=6)
for seed in [int(x) for x in sys.argv[1:]]:
    lat, lon, dep, comps, stype, base, cands, cfg = t.random_case(seed)
    print("seed", seed, stype, cfg, comps)
    db = getattr(sc, cfg["db"])()
    g, o, w = Engine(0), OracleEngine(), OracleEngine(wide=True)
    for e in (g, o, w):
        sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
    if os.environ.get('DISC'):
        o.discretize_source(stype, base); g.discretize_source(stype, base)
    o.set_source_params(stype, base)
    refs = sc.set_refs_from(o, [g, o, w], [len(c) for c in comps])
    for e in (g, o, w):
        e.set_misfit_method(cfg["norm"]); e.set_synthetics_factor(cfg["factor"])
        if cfg["norm"].startswith("floating"): e.set_floating_shiftrange(-0.4, 0.3)
        if cfg["taper"]:
            for ir in range(1, len(comps) + 1): e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
        if cfg["filt"]: e.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
        if cfg["disable"]: e.switch_receiver(cfg["disable"], False)
    if cfg.get("autoshift"):
        g.set_source_params(stype, base); w.set_source_params(stype, base)
        a, b, c = g.autoshift_ref_seismogram(0, -0.3, 0.4), o.autoshift_ref_seismogram(0, -0.3, 0.4), w.autoshift_ref_seismogram(0, -0.3, 0.4)
        print("autoshift", a, b, c)
    mg, sg = g.eval_sources(stype, cands); mo, so = o.eval_sources(stype, cands); mw, sw = w.eval_sources(stype, cands)
    print("status", sg, so, sw)
    d = np.abs(mg - mw) / np.maximum(np.abs(mw), 0.1 * np.abs(mw[..., 1:2]))
    i = np.unravel_index(np.argmax(d), d.shape)
    print("worst vs wide at", i, "gpu", mg[i[0], i[1]], "fp32", mo[i[0], i[1]], "wide", mw[i[0], i[1]])
    for ir in range(1, len(comps) + 1):
        for ic in range(1, len(comps[ir - 1]) + 1):
            try:
                print(" spans rcv", ir, ic, "oracle", o.get_probe_spans(ir, ic), "wide", w.get_probe_spans(ir, ic))
            except Exception as ex:
                print(" spans n/a", ex); break
