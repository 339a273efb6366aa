import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import scenario as sc
import test_parity_gpu as T
from oracle_lib import OracleEngine
from kiwi_b200 import Engine
lat, lon, dep = sc.small_receivers(6)
g, o, w = Engine(0), OracleEngine(), OracleEngine(wide=True)
for e in (g, o, w):
    sc.setup(e, sc.small_db(), lat, lon, dep, T.COMPS6)
o.eval_sources("bilateral", sc.BILAT_SMALL)
sc.set_refs_from(o, [g, o, w], [len(c) for c in T.COMPS6])
for e in (g, o, w):
    e.set_misfit_method("l2norm"); e.set_misfit_filter(*T.FILTER)
p = T._candidates()
mg, _ = g.eval_sources("bilateral", p); mo, _ = o.eval_sources("bilateral", p); mw, _ = w.eval_sources("bilateral", p)
tol = T.misfit_tol(mw)
for name, a, b in (("g-o", mg, mo), ("g-w", mg, mw), ("o-w", mo, mw)):
    r = np.abs(a - b) / tol
    i = np.unravel_index(np.argmax(r), r.shape)
    print(name, r.max(), i, a[i], b[i], mw[i[0], i[1], :])
