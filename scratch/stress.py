"""Stress of the first evaluation of fresh engine contexts under concurrent load (run several copies at once on one GPU):
    for i in 1 2 3 4; do python scratch/stress.py 60 340 161 & done; wait
Every repetition builds a new engine, evaluates a seeded random case (tests/test_random_parity_gpu.py) twice and compares with the
oracle; a mismatch of the FIRST evaluation only is how the stream-ordering bug of profiles/r01_sanitizer.md showed."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenario as sc
import test_random_parity_gpu as t
from oracle_lib import OracleEngine
from kiwi_b200 import Engine
seeds = [int(x) for x in sys.argv[2:]]
nrep = int(sys.argv[1])
bad = 0
for rep in range(nrep):
    for seed in seeds:
        lat, lon, dep, comps, stype, base, cands, cfg = t.random_case(seed)
        db = getattr(sc, cfg["db"])()
        g, o = Engine(0), OracleEngine()
        for e in (g, o):
            sc.setup(e, db, lat, lon, dep, comps, interpolation=cfg["interp"], effective_dt=cfg["eff_dt"], under=cfg["under"])
        g.discretize_source(stype, base)
        o.set_source_params(stype, base)
        refs = sc.set_refs_from(o, [g, o], [len(c) for c in comps])
        for e in (g, o):
            e.set_misfit_method(cfg["norm"]); e.set_synthetics_factor(cfg["factor"])
            if cfg["taper"]:
                for ir in range(1, len(comps) + 1): e.set_misfit_taper(ir, [0.8, 1.5, 4.5, 5.5], [0, 1, 1, 0])
            if cfg["filt"]: e.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
            if cfg["disable"]: e.switch_receiver(cfg["disable"], False)
        mg, sg = g.eval_sources(stype, cands); mo, so = o.eval_sources(stype, cands)
        mg2, sg2 = g.eval_sources(stype, cands)
        d = np.abs(mg - mo).max() / np.abs(mo).max()
        d2 = np.abs(mg - mg2).max() / np.abs(mo).max()
        if d > 1e-4 or d2 > 0:
            bad += 1
            print("MISMATCH pid", os.getpid(), "rep", rep, "seed", seed, "gpu-vs-oracle %.3g gpu-vs-gpu(again) %.3g" % (d, d2), "gpu", mg[0, 0], "again", mg2[0, 0], "oracle", mo[0, 0], flush=True)
print("pid", os.getpid(), "done, mismatches", bad)
