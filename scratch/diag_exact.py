"""Reference-order synthesis against the fp32 oracle on the small scenario and at C3 size."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scenario as sc
from test_parity_gpu import COMPS6, EIK, engines
from kiwi_b200 import synthetic

def compare(g, o, nrcv, ncomps, label):
    worst = 0.0; nsame = ntot = 0
    for ir in range(1, nrcv + 1):
        for ic in range(1, ncomps[ir - 1] + 1):
            (fg, dg), (fo, do) = g.get_seismogram(ir, ic), o.get_seismogram(ir, ic)
            assert (fg, dg.size) == (fo, do.size), (label, ir, ic, fg, dg.size, fo, do.size)
            peak = float(np.abs(do).max())
            worst = max(worst, float(np.abs(dg - do).max()) / peak)
            nsame += int((dg.view(np.uint32) == do.view(np.uint32)).sum()); ntot += dg.size
    print("%-40s max |gpu - oracle| / peak = %.3e   bit-identical samples %d of %d" % (label, worst, nsame, ntot), flush=True)

ncomps = [len(c) for c in COMPS6]
EIK0 = EIK.copy(); EIK0[14] = 0.0
for stype, p in (("bilateral", sc.BILAT_SMALL), ("moment_tensor", sc.MT_SMALL), ("eikonal", EIK0)):
    g, o = engines(sc.small_db(), COMPS6)
    o.eval_sources(stype, p)
    for mode in (0, 1):
        g.set_accumulation(mode)
        g.set_source_params(stype, p)
        compare(g, o, 6, ncomps, "%s small, mode %d" % (stype, mode))
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from test_fullsize_parity_gpu import setup_pair
    g, o, w = setup_pair("c3", nrcv=int(sys.argv[1]))
    o.eval_sources("bilateral", synthetic.IZMIT)
    for mode in (0, 1):
        g.set_accumulation(mode)
        t0 = time.perf_counter(); g.set_source_params("bilateral", synthetic.IZMIT); g.get_seismogram(1, 1); t1 = time.perf_counter()
        compare(g, o, w["nrcv"], [3] * w["nrcv"], "C3 (%d receivers) mode %d, %.2f s" % (w["nrcv"], mode, t1 - t0))
