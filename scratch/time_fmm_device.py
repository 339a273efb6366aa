"""Throughput of the device fast-marching solver against the host solver: njobs grids of n x n nodes (uniform layered field)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiwi_b200 import engine

def field(n, seed):
    sp = np.repeat(np.array([2400.0, 3100.0, 3600.0], np.float32)[np.minimum(np.arange(n) * 3 // n, 2)][:, None], n, 1).copy()
    yy, xx = np.mgrid[0:n, 0:n]
    sp[(xx - n / 2.0) ** 2 + (yy - n / 2.0) ** 2 > (0.5 * n) ** 2] = 1200.0
    return sp

for n in (400, 700):
    sp = field(n, 0)
    t0 = time.perf_counter(); host = engine.eikonal_fmm(sp, (0, 0), (25, 25), (n * 12.5 - 3000, n * 12.5 + 1500)); th = time.perf_counter() - t0
    print("host  %dx%d: %.1f ms (%.0f ns/node)" % (n, n, th * 1e3, th * 1e9 / n / n), flush=True)
    for njobs in (1, 32, 148, 592, 1036, 2072):
        if n == 700 and njobs > 1036:
            continue
        dev, ms = engine.eikonal_fmm_device([sp] * njobs, [(0, 0)] * njobs, [(25, 25)] * njobs, [(n * 12.5 - 3000, n * 12.5 + 1500)] * njobs)
        ok = all(np.array_equal(d.view(np.uint32), host.view(np.uint32)) for d in dev[:3] + dev[-1:])
        print("device %dx%d x %4d jobs: %.1f ms  -> %.0f ns/node/job, %.1f ns/node aggregate, %.1f solves/s, bit-exact %s"
              % (n, n, njobs, ms, ms * 1e6 / n / n, ms * 1e6 / n / n / njobs, njobs / (ms * 1e-3), ok), flush=True)
