import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kiwi_b200 import engine
n = 300
sp = np.repeat(np.array([2400.0, 3100.0, 3600.0], np.float32)[np.minimum(np.arange(n) * 3 // n, 2)][:, None], n, 1).copy()
yy, xx = np.mgrid[0:n, 0:n]
sp[(xx - n / 2.0) ** 2 + (yy - n / 2.0) ** 2 > (0.5 * n) ** 2] = 1200.0
njobs = 296
dev, ms = engine.eikonal_fmm_device([sp] * njobs, [(0, 0)] * njobs, [(25, 25)] * njobs, [(n * 12.5 - 1000, n * 12.5 + 500)] * njobs)
print(ms)
