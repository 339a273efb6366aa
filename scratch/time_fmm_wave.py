"""Device fast-marching solver at one and two waves' worth of resident solves, with the 16 KB and the 8 KB heap (KIWI_EIKONAL_HEAP)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiwi_b200 import engine
from time_fmm_device_field import field

n = int(sys.argv[1]); jobs = [int(a) for a in sys.argv[2:]]
sp = field(n, 0)
host = engine.eikonal_fmm(sp, (0, 0), (25, 25), (n * 12.5 - 3000, n * 12.5 + 1500))
for njobs in jobs:
    dev, ms = engine.eikonal_fmm_device([sp] * njobs, [(0, 0)] * njobs, [(25, 25)] * njobs, [(n * 12.5 - 3000, n * 12.5 + 1500)] * njobs)
    ok = all(np.array_equal(d.view(np.uint32), host.view(np.uint32)) for d in dev[:3] + dev[-1:])
    print("heap %s  %dx%d x %4d jobs: %.1f ms -> %.0f ns/node/job, %.2f ns/node aggregate, %.0f solves/s, bit-exact %s"
          % (os.environ.get("KIWI_EIKONAL_HEAP", "auto"), n, n, njobs, ms, ms * 1e6 / n / n, ms * 1e6 / n / n / njobs, njobs / (ms * 1e-3), ok), flush=True)
