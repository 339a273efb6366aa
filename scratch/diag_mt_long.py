import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scenario as sc
from test_limits_gpu import long_trace_db
from kiwi_b200 import Engine, synthetic
from oracle_lib import OracleEngine
db = long_trace_db()
lat, lon, dep = sc.small_receivers(4, dmin=26e3, dmax=36e3)
comps = ["ned", "ar", "d", "neu"]
g, o = Engine(0), OracleEngine()
for e in (g, o):
    sc.setup(e, db, lat, lon, dep, comps, effective_dt=1.0)
base = np.array([1.0, 500, -800, 2500, 1e18, -0.4e18, -0.6e18, 0.3e18, 0.2e18, -0.5e18, 2.0], dtype=np.float32)
o.eval_sources("moment_tensor", base)
sc.set_refs_from(o, [g, o], [len(c) for c in comps], dt=0.5)
mts = synthetic.fibonacci_moment_tensors(40) * 1e18
p = np.tile(base, (2 * 40, 1))
p[:, 4:10] = np.concatenate([mts, mts])
p[40:, 1] += 700; p[40:, 3] += 500
mo, so = o.eval_sources("moment_tensor", p)
for mode in (1, 2, 0):
    g.set_mt_grid(mode if mode != 1 else True)
    mg, sg = g.eval_sources("moment_tensor", p)
    bad = np.argwhere(~np.isfinite(mg))
    print("mode", mode, "launches", g.last_timing()["launches"], "status nonzero", int((sg != 0).sum()), "nonfinite entries", bad.shape[0], bad[:12].tolist())
    ok = np.isfinite(mg)
    print("   max dev where finite", float(np.max(np.abs(mg - mo)[ok] / np.maximum(np.abs(mo), 0.1 * np.abs(mo[..., 1:2]))[ok])))
