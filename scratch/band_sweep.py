"""One-off experiment: C5 step time as a function of KIWI_SYNTH_BANDS (depth-band launches of k_synth)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import kiwi_b200  # noqa: E402
from kiwi_b200 import synthetic  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
settings = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,15,8,5,3,1").split(",")]
w = dict(bench.WORKLOADS[wl])
db = synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"])
rlat, rlon, rdep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
eng = kiwi_b200.Engine(0)
bench.configure(eng, db, w, rlat, rlon, rdep)


class A:
    workload = wl
    batch = w["batch"]


stype, allc, base = bench.candidates(w, max(A.batch, 32))
eng.set_source_params(stype, base)
bench.set_references(eng, [eng], w["nrcv"], db.meta()["dt"])
mine = np.ascontiguousarray(allc[:A.batch])
ref = None
for nb in settings:
    os.environ["KIWI_SYNTH_BANDS"] = str(nb)
    for _ in range(2):
        mis, st = eng.eval_sources(stype, mine)
    ts = []
    for _ in range(4):
        mis, st = eng.eval_sources(stype, mine)
        ts.append(eng.last_timing())
    if ref is None:
        ref = mis.copy()
    dev = float(np.max(np.abs(mis - ref) / np.maximum(np.abs(ref), 0.1 * np.abs(ref[..., 1:2]))))
    print(json.dumps({"bands": nb, "synth_ms": float(np.mean([t["synthesis_ms"] for t in ts])), "total_ms": float(np.mean([t["total_ms"] for t in ts])),
                      "evals_per_s": A.batch / (np.mean([t["total_ms"] for t in ts]) * 1e-3), "dev_vs_first": dev, "status_any": bool(st.any())}), flush=True)
