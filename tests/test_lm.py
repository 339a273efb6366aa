"""Levenberg-Marquardt.  The reference links MINPACK's lmdif in single precision (sminpack/)
and has no tests for it, so the restatements are pinned on MINPACK's own published test problems (More, Garbow,
Hillstrom, "Testing unconstrained optimization software", ACM TOMS 7, 1981; final residual norms as listed in the
MINPACK-1 test output).  The product's batched lmdif (kiwi_lmdif_batched: all Jacobian columns in one call) must
then agree bit for bit with the sequential restatement in oracle/ko_lm.hpp."""
import numpy as np
import pytest

from kiwi_b200 import lmdif_batched
from oracle_lib import oracle_enorm, oracle_lmdif

F = np.float32


def linear_full_rank(n, m):
    def f(x):
        s = F(0)
        for v in x:
            s = F(s + v)
        temp = F(F(2) * s / F(m) + F(1))
        out = np.full(m, -temp, F)
        out[:n] = x - temp
        return out
    return f, np.ones(n, F), m, 2.236068 if (n, m) == (5, 10) else None


def rosenbrock():
    def f(x):
        return np.array([F(10) * (x[1] - x[0] * x[0]), F(1) - x[0]], F)
    return f, np.array([-1.2, 1.0], F), 2, 0.0


def powell_singular():
    def f(x):
        return np.array([x[0] + F(10) * x[1], np.sqrt(F(5)) * (x[2] - x[3]), (x[1] - F(2) * x[2]) ** 2, np.sqrt(F(10)) * (x[0] - x[3]) ** 2], F)
    return f, np.array([3, -1, 0, 1], F), 4, 0.0


def freudenstein_roth():
    def f(x):
        return np.array([-F(13) + x[0] + ((F(5) - x[1]) * x[1] - F(2)) * x[1], -F(29) + x[0] + ((F(1) + x[1]) * x[1] - F(14)) * x[1]], F)
    return f, np.array([0.5, -2.0], F), 2, 6.998875


def bard():
    y = np.array([0.14, 0.18, 0.22, 0.25, 0.29, 0.32, 0.35, 0.39, 0.37, 0.58, 0.73, 0.96, 1.34, 2.10, 4.39], F)

    def f(x):
        out = np.zeros(15, F)
        for i in range(15):
            t1 = F(i + 1); t2 = F(15 - i); t3 = t1 if t1 < t2 else t2
            out[i] = y[i] - (x[0] + t1 / (x[1] * t2 + x[2] * t3))
        return out
    return f, np.ones(3, F), 15, 0.09063596


def kowalik_osborne():
    v = np.array([4.0, 2.0, 1.0, 0.5, 0.25, 0.167, 0.125, 0.1, 0.0833, 0.0714, 0.0625], F)
    y = np.array([0.1957, 0.1947, 0.1735, 0.1600, 0.0844, 0.0627, 0.0456, 0.0342, 0.0323, 0.0235, 0.0246], F)

    def f(x):
        t1 = v * (v + x[1]); t2 = v * (v + x[2]) + x[3]
        return (y - x[0] * t1 / t2).astype(F)
    return f, np.array([0.25, 0.39, 0.415, 0.39], F), 11, 0.01753584


def brown_dennis():
    def f(x):
        out = np.zeros(20, F)
        for i in range(20):
            t = F(i + 1) / F(5)
            t1 = x[0] + t * x[1] - np.exp(t); t2 = x[2] + np.sin(t) * x[3] - np.cos(t)
            out[i] = t1 * t1 + t2 * t2
        return out
    return f, np.array([25, 5, -5, -1], F), 20, 292.9543


def box3d():
    def f(x):
        t = (np.arange(1, 11, dtype=F) / F(10)).astype(F)
        return (np.exp(-t * x[0]) - np.exp(-t * x[1]) + (np.exp(-F(10) * t) - np.exp(-t)) * x[2]).astype(F)
    return f, np.array([0, 10, 20], F), 10, 0.0


PROBLEMS = {"linear_full_rank": lambda: linear_full_rank(5, 10), "rosenbrock": rosenbrock, "powell_singular": powell_singular,
            "freudenstein_roth": freudenstein_roth, "bard": bard, "kowalik_osborne": kowalik_osborne, "brown_dennis": brown_dennis, "box3d": box3d}


def _f32fn(f):
    def g(x):
        with np.errstate(all="ignore"):
            return np.asarray(f(np.asarray(x, F)), F)
    return g


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_minpack_problems_known_answers_and_batched_equals_sequential(name):
    f, x0, m, fnorm_known = PROBLEMS[name]()
    f = _f32fn(f)
    xo, fo, info_o, nfev_o = oracle_lmdif(f, x0, m)
    assert 1 <= info_o <= 8, info_o
    fn = float(np.sqrt(np.sum(fo.astype(np.float64) ** 2)))
    if fnorm_known == 0.0:
        assert fn < 2e-3, fn
    else:
        assert abs(fn - fnorm_known) <= 2e-3 * fnorm_known, (fn, fnorm_known)
    xb, fb, info_b, nfev_b = lmdif_batched(lambda xs: np.stack([f(x) for x in xs]), x0, m)
    assert (info_b, nfev_b) == (info_o, nfev_o)
    assert np.array_equal(xb.view(np.uint32), xo.view(np.uint32)) and np.array_equal(fb.view(np.uint32), fo.view(np.uint32))


def test_engine_settings_mode2_and_in_place_clipping():
    """the settings minimize_lm uses (mode 2, diag 1, factor 0.01, gtol 0; minimizer_engine.f90:783-797) and a forward step that
    clips its argument in place with a penalty factor (:829-848)"""
    f0, x0, m, _ = bard()
    f0 = _f32fn(f0)
    lo, hi = np.array([0.0, 0.5, 0.5], F), np.array([0.05, 5.0, 5.0], F)   # the optimum of x[0] (0.0824) lies outside

    def clip_eval(x):
        pen = F(0)
        for i in range(3):
            if x[i] < lo[i]:
                pen = F(pen + abs(x[i] - lo[i]) / abs(hi[i] - lo[i])); x[i] = lo[i]
            if x[i] > hi[i]:
                pen = F(pen + abs(x[i] - hi[i]) / abs(hi[i] - lo[i])); x[i] = hi[i]
        return (f0(x) * (F(1) + pen)).astype(F)

    def batched(xs):
        return np.stack([clip_eval(xs[i]) for i in range(xs.shape[0])])   # rows of xs are views: clipped in place

    kw = dict(gtol=0.0, maxfev=500 * 4, epsfcn=0.0, diag=np.ones(3, F), mode=2, factor=0.01)
    xo, fo, info_o, nfev_o = oracle_lmdif(clip_eval, x0, m, **kw)
    xb, fb, info_b, nfev_b = lmdif_batched(batched, x0, m, **kw)
    assert (info_b, nfev_b) == (info_o, nfev_o) and 1 <= info_o <= 8
    assert np.array_equal(xb.view(np.uint32), xo.view(np.uint32)) and np.array_equal(fb.view(np.uint32), fo.view(np.uint32))
    assert np.all(xo >= lo) and np.all(xo <= hi) and abs(xo[0] - hi[0]) < 1e-6


def test_failure_of_the_function_stops_with_iflag():
    f, x0, m, _ = rosenbrock()
    f = _f32fn(f)
    calls = {"n": 0}

    def failing(x):
        calls["n"] += 1
        return None if calls["n"] == 6 else f(x)

    xo, fo, info_o, nfev_o = oracle_lmdif(failing, x0, m)
    assert info_o == -2
    calls["n"] = 0

    def failing_batched(xs):
        out = []
        for x in xs:
            r = failing(x)
            if r is None:
                break
            out.append(r)
        return np.stack(out) if out else None

    xb, fb, info_b, nfev_b = lmdif_batched(failing_batched, x0, m)
    assert info_b == -2 and nfev_b == nfev_o
    assert np.array_equal(xb.view(np.uint32), xo.view(np.uint32))


def test_improper_input_and_enorm():
    f, x0, m, _ = rosenbrock()
    assert lmdif_batched(lambda xs: np.stack([_f32fn(f)(x) for x in xs]), x0, 1)[2] == 0      # m < n
    assert oracle_lmdif(_f32fn(f), x0, 1)[2] == 0
    assert oracle_enorm([3.0, 4.0]) == 5.0
    assert abs(oracle_enorm([3e-25, 4e-25]) / 5e-25 - 1) < 1e-6        # below rdwarf: no underflow
    assert abs(oracle_enorm([3e22, 4e22]) / 5e22 - 1) < 1e-6            # above rgiant/n: no overflow
    assert oracle_enorm([0.0, 0.0, 0.0]) == 0.0
