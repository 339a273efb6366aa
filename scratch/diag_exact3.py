import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiwi_b200 import synthetic
from test_fullsize_parity_gpu import setup_pair
g, o, w = setup_pair("c3", nrcv=12)
o.record_indices(True)
o.eval_sources("bilateral", synthetic.IZMIT)
g.set_accumulation(1)
g.set_source_params("bilateral", synthetic.IZMIT)
for ir in range(1, 13):
    ig, io = g.get_indices(ir), o.get_indices(ir)
    dd = ig["dix"].view(np.uint32) != io["dix"].view(np.uint32)
    dz = ig["diz"].view(np.uint32) != io["diz"].view(np.uint32)
    near = ig["near"] != 0
    k = np.flatnonzero(dd)
    print("rcv %2d: dix differ %d (groups %d), diz differ %d, flagged %d, differ&flagged %d; first: %s" % (
        ir, dd.sum(), len(set((k // 7).tolist())), dz.sum(), near.sum(), (dd & near).sum(),
        [(int(i), float(ig["dix"][i]), float(io["dix"][i]), "%.6f" % io["dist"][i]) for i in k[:2]]), flush=True)
