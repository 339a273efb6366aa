import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from kiwi_b200 import synthetic
from test_fullsize_parity_gpu import setup_pair, bench_l
for name in ("c3", "c5"):
    g, o, w = setup_pair(name)
    g.set_source_params("bilateral", synthetic.IZMIT)
    bench.set_references(g, [g, o], w["nrcv"], bench_l().meta()["dt"])
    p = synthetic.bilateral_sweep(32)[:2]
    for mode in (0, 1):
        g.set_accumulation(mode)
        mg, sg = g.eval_sources("bilateral", p)
        mg, sg = g.eval_sources("bilateral", p)      # (second call: buffers and worker threads exist)
        t = g.last_timing()
        print(name, "mode", mode, "per candidate: synthesis ms", t["synthesis_ms"] / 2, "geometry ms", t["geometry_ms"] / 2, "total ms", t["total_ms"] / 2, flush=True)
        if name == "c3":
            mo, so = o.eval_sources("bilateral", p)
            print("   misfits vs fp32 oracle: max rel dev", float(np.max(np.abs(mg - mo) / np.maximum(np.abs(mo), 1e-30))), " rel to norm factor", float(np.max(np.abs(mg - mo) / np.abs(mo[..., 1:2]))), flush=True)
