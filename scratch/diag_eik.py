import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from kiwi_b200 import synthetic, Gfdb
from oracle_lib import OracleEngine
w = dict(bench.WORKLOADS["c4"])
db = Gfdb.create(20, 20, 10, 0.1, 100.0, 200.0, 100.0, 0.0).build_ahfull(2700.0, 6000.0, 3464.0)
rlat, rlon, rdep = synthetic.receivers(2, (30.0, 70.0), 500, 1500)
o = OracleEngine(threads=1)
bench.configure(o, db, dict(w, taper=None, filter=None), rlat, rlon, rdep)
stype, allc, base = bench.candidates(w, 32)
for i in (0, 1, 3):
    t = time.time(); tab, grid, n = o.discretize_source(stype, allc[i]); print(i, allc[i][10], grid, n, "%.1f ms" % (1e3 * (time.time() - t)))
