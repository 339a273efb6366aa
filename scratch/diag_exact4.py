import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiwi_b200 import synthetic
from test_fullsize_parity_gpu import setup_pair
g, o, w = setup_pair("c3", nrcv=60)
o.eval_sources("bilateral", synthetic.IZMIT)
g.set_accumulation(1)
g.set_source_params("bilateral", synthetic.IZMIT)
tot = [0, 0, 0]; diff = [0, 0, 0]; rc_bad = []
pos = []
for ir in range(1, 61):
    nb = 0
    for ic in range(1, 4):
        (fg, dg), (fo, do) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic)
        ne = dg.view(np.uint32) != do.view(np.uint32)
        tot[ic - 1] += dg.size; diff[ic - 1] += int(ne.sum()); nb += int(ne.sum())
        if ne.any():
            k = np.flatnonzero(ne)
            pos.append((ir, ic, int(k[0]), int(k[-1]), dg.size, int(ne.sum()), int(np.abs(dg.view(np.int32)[k] - do.view(np.int32)[k]).max())))
    rc_bad.append(nb)
print("samples", tot, "non-identical", diff)
print("receivers with no difference:", sum(1 for v in rc_bad if v == 0), "of 60;  per receiver:", rc_bad)
for p in pos[:25]:
    print("rcv %d comp %d: first %d last %d of %d, count %d, max ulp distance %d" % p)
