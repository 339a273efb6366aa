"""The host fast-marching solver of the eikonal sources (kiwi_b200/csrc/source_eikonal_host.cpp) without a GPU: the reference's own
known-answer test (test_eikonal.f90:26-56) and bit-exactness against the restatement on random speed fields -- the solver uses
tentative neighbour values, so its result depends on the exact pop order of the reference's heap."""
import ctypes as C
import time

import numpy as np

import oracle_lib as ol
from kiwi_b200 import engine

fp = C.POINTER(C.c_float)


def oracle_fmm(speed, origin, delta, ip):
    sp = np.ascontiguousarray(speed, np.float32)
    ny, nx = sp.shape
    times = np.zeros_like(sp)
    o, d, p = (np.asarray(v, np.float32) for v in (origin, delta, ip))
    ol.lib().oracle_eikonal_fmm(C.c_int(nx), C.c_int(ny), sp.ctypes.data_as(fp), o.ctypes.data_as(fp), d.ctypes.data_as(fp), p.ctypes.data_as(fp),
                                times.ctypes.data_as(fp))
    return times


def test_reference_known_answer():
    nx, ny = 500, 1000
    delta = (50.0 / nx, 50.0 / ny)
    t = engine.eikonal_fmm(np.full((ny, nx), 2.0, np.float32), (0.0, 0.0), delta, (0.0, 25.0))
    eps = max(delta) / 2.0
    assert abs(t[0, 0] - 12.5) < eps and abs(t[ny - 1, 0] - 12.5) < eps            # times(1,1), times(1,ny)
    assert abs(t[0, nx - 1] - 27.95) < eps and abs(t[ny - 1, nx - 1] - 27.95) < eps  # times(nx,1), times(nx,ny)


def test_bit_exact_against_the_restatement_on_random_fields():
    rng = np.random.default_rng(11)
    for nx, ny in ((1, 1), (1, 17), (23, 1), (40, 31), (120, 90), (300, 200)):
        for kind in range(3):
            if kind == 0:
                sp = np.full((ny, nx), 2800.0, np.float32)                         # uniform: many exactly equal keys in the heap
            elif kind == 1:
                sp = rng.uniform(1500.0, 3500.0, (ny, nx)).astype(np.float32)
            else:                                                                    # a slow inclusion with a sharp edge and an invalid rim
                sp = np.full((ny, nx), 3000.0, np.float32)
                sp[ny // 3: ny // 2 + 1, nx // 4: nx // 2 + 1] = 900.0
                sp[0, :] = 450.0
            delta = (25.0, float(rng.choice([25.0, 40.0])))
            ip = (float(rng.uniform(-10, nx * delta[0] + 10)), float(rng.uniform(-10, ny * delta[1] + 10)))
            a = engine.eikonal_fmm(sp, (0.0, 0.0), delta, ip)
            b = oracle_fmm(sp, (0.0, 0.0), delta, ip)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nx, ny, kind)


def test_speed_of_the_solver_is_reported():
    n = 600
    sp = np.full((n, n), 2800.0, np.float32)
    t0 = time.perf_counter()
    engine.eikonal_fmm(sp, (0.0, 0.0), (25.0, 25.0), (7500.0, 7500.0))
    dt = time.perf_counter() - t0
    print("host fast marching: %.0f ns per node" % (1e9 * dt / n / n))
    assert dt < 30.0
