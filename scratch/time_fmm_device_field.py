import numpy as np


def field(n, seed):
    sp = np.repeat(np.array([2400.0, 3100.0, 3600.0], np.float32)[np.minimum(np.arange(n) * 3 // n, 2)][:, None], n, 1).copy()
    yy, xx = np.mgrid[0:n, 0:n]
    sp[(xx - n / 2.0) ** 2 + (yy - n / 2.0) ** 2 > (0.5 * n) ** 2] = 1200.0
    return sp
