import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from kiwi_b200 import synthetic
from oracle_lib import OracleEngine
w = dict(bench.WORKLOADS["c4"]); w.update(nx=1000, nz=80, dz=400.0, nrcv=6, dmin=45e3, dmax=55e3)
t=time.time(); db = synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"]); print("db", time.time()-t)
rlat, rlon, rdep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
stype, allc, base = bench.candidates(w, 32)
res = {}
for name, wide in (("f32", False), ("wide", True)):
    o = OracleEngine(threads=8, wide=wide)
    bench.configure(o, db, w, rlat, rlon, rdep)
    if name == "f32":
        o.eval_sources(stype, base)
        bench.set_references(o, [o], w["nrcv"], db.meta()["dt"])
    else:
        bench.copy_references(None, o, w["nrcv"], db.meta()["dt"])
    m, s = o.eval_sources(stype, allc[:2])
    res[name] = m
    print(name, s, m[0, :6])
a, b = res["f32"], res["wide"]
print(np.abs(a - b)[0, :9], np.max(np.abs(a - b) / np.maximum(np.abs(b), 0.1 * np.abs(b[..., 1:2]))))
for norm in ("l2norm", "ampspec_l2norm"):
    out = []
    for wide in (False, True):
        o = OracleEngine(threads=8, wide=wide)
        ww = dict(w); ww["norm"] = norm
        bench.configure(o, db, ww, rlat, rlon, rdep)
        bench.copy_references(None, o, w["nrcv"], db.meta()["dt"])
        m, s = o.eval_sources(stype, allc[:2]); out.append(m)
    print(norm, np.max(np.abs(out[0] - out[1]) / np.maximum(np.abs(out[1]), 0.1 * np.abs(out[1][..., 1:2]))))
