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
    nd = int((ig["ix"] != io["ix"]).sum()); nn = int((ig["near"] != 0).sum())
    ddix = float(np.abs(ig["dix"] - io["dix"]).max()) if "dix" in ig else -1
    out = []
    for ic in range(1, 4):
        (fg, dg), (fo, do) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic)
        peak = np.abs(do).max(); d = np.abs(dg - do) / peak
        same = (dg.view(np.uint32) == do.view(np.uint32))
        k = int(np.argmax(d))
        out.append("c%d max %.1e at %d/%d same %.2f" % (ic, d.max(), k, dg.size, same.mean()))
    print("rcv %2d: ix differ %d, flagged %d, max|ddix| %.1e | %s" % (ir, nd, nn, ddix, " | ".join(out)), flush=True)
