"""The device fast-marching solver (kiwi_b200/csrc/eikonal.cu: one warp per grid replaying heap.f90) against the host solver, bit for
bit: uniform fields (many exactly equal keys in the heap), layered fields with a slow rim (the eikonal sources' speed fields), random
fields, degenerate grids, heaps that outgrow the shared-memory part."""
import numpy as np
import pytest

from kiwi_b200 import engine

pytestmark = pytest.mark.gpu


def fields():
    rng = np.random.default_rng(5)
    out = []
    for nx, ny in ((1, 1), (1, 17), (23, 1), (2, 2), (40, 31), (120, 90), (301, 200)):
        for kind in range(3):
            if kind == 0:
                sp = np.full((ny, nx), 2800.0, np.float32)
            elif kind == 1:   # layers along y, half-speed outside a disc: what psm_make_eikonal_grid hands to the solver
                sp = np.repeat(np.array([2400.0, 3100.0, 3600.0], np.float32)[np.minimum(np.arange(ny) * 3 // max(ny, 1), 2)][:, None], nx, 1).copy()
                yy, xx = np.mgrid[0:ny, 0:nx]
                sp[(xx - nx / 2.0) ** 2 + (yy - ny / 2.0) ** 2 > (0.45 * max(nx, ny)) ** 2] = 1200.0
            else:
                sp = rng.uniform(500.0, 5000.0, (ny, nx)).astype(np.float32)
            delta = (25.0, 25.0) if kind < 2 else (float(rng.uniform(10, 40)), float(rng.uniform(10, 40)))
            ip = (float(rng.uniform(-10, nx * delta[0] + 10)), float(rng.uniform(-10, ny * delta[1] + 10)))
            out.append((sp, (0.0, 0.0), delta, ip))
    return out


def test_bit_exact_against_the_host_solver():
    cases = fields()
    dev, ms = engine.eikonal_fmm_device([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases])
    for (sp, o, d, ip), t in zip(cases, dev):
        host = engine.eikonal_fmm(sp, o, d, ip)
        assert np.array_equal(host.view(np.uint32), t.view(np.uint32)), (sp.shape, d, ip, float(np.abs(host - t).max()))


@pytest.fixture(params=["2040", "980"])
def heap_build(request, monkeypatch):
    """both builds of the solver: 2040 heap entries in shared memory (13 solves per SM, batches of one wave) and 980 (26 per SM, larger
    batches); KIWI_EIKONAL_HEAP forces one whatever the batch size"""
    monkeypatch.setenv("KIWI_EIKONAL_HEAP", request.param)
    return request.param


def test_both_heap_builds_bit_exact(heap_build):
    cases = fields()
    dev, ms = engine.eikonal_fmm_device([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases])
    for (sp, o, d, ip), t in zip(cases, dev):
        host = engine.eikonal_fmm(sp, o, d, ip)
        assert np.array_equal(host.view(np.uint32), t.view(np.uint32)), (heap_build, sp.shape)


def test_front_around_the_shared_memory_boundary(heap_build):
    """layered 500 x 500 disc field (what the eikonal sources solve): the front grows through 980 entries and shrinks back, so entries
    move between the shared-memory part and the global overflow in both directions"""
    n = 500
    sp = np.repeat(np.array([2400.0, 3100.0, 3600.0], np.float32)[np.minimum(np.arange(n) * 3 // n, 2)][:, None], n, 1).copy()
    yy, xx = np.mgrid[0:n, 0:n]
    sp[(xx - n / 2.0) ** 2 + (yy - n / 2.0) ** 2 > (0.5 * n) ** 2] = 1200.0
    ip = (n * 12.5 - 3000.0, n * 12.5 + 1500.0)
    dev, ms = engine.eikonal_fmm_device([sp, sp[:, ::-1].copy()], [(0.0, 0.0)] * 2, [(25.0, 25.0)] * 2, [ip] * 2)
    for field, t in zip((sp, sp[:, ::-1].copy()), dev):
        host = engine.eikonal_fmm(field, (0.0, 0.0), (25.0, 25.0), ip)
        assert np.array_equal(host.view(np.uint32), t.view(np.uint32))


def test_heap_larger_than_its_shared_memory_part(heap_build):
    """a thin, long, fast channel in a slow field makes the front (the heap) longer than 4096 entries"""
    ny, nx = 900, 1200
    sp = np.full((ny, nx), 400.0, np.float32)
    sp[::2, :] = 6000.0          # every other row fast: the front runs along all of them at once
    sp[:, 0] = 6000.0
    dev, ms = engine.eikonal_fmm_device([sp], [(0.0, 0.0)], [(25.0, 25.0)], [(0.0, 0.0)])
    host = engine.eikonal_fmm(sp, (0.0, 0.0), (25.0, 25.0), (0.0, 0.0))
    assert np.array_equal(host.view(np.uint32), dev[0].view(np.uint32))


def test_batch_of_eikonal_sources_solved_on_the_device_equals_the_host_path(heap_build):
    """the engine's device path for large batches (kiwi_set_eikonal_device): sub-source tables and misfits identical to the host path's"""
    import scenario as sc
    from test_parity_gpu import COMPS6, EIK, engines
    g, o = engines(sc.small_db(), COMPS6)
    o.eval_sources("eikonal", EIK)
    sc.set_refs_from(o, [g], [len(c) for c in COMPS6])
    n = 40
    p = np.tile(EIK, (n, 1))
    i = np.arange(n)
    p[:, 10] = np.linspace(600.0, 1800.0, 5)[i % 5]            # bord radius
    p[:, 13] = np.linspace(0.6, 0.9, 4)[(i // 5) % 4]          # relative rupture velocity
    p[:, 11] = np.linspace(-300.0, 300.0, 2)[(i // 20) % 2]    # nucleation point
    p[7, 11] = 5000.0                                          # outside the rupture area: this candidate fails, the others do not care
    g.set_eikonal_device(0)
    mh, sh = g.eval_sources("eikonal", p)
    th = [g.discretize_source("eikonal", q) for q in p[:3]]
    g.set_eikonal_device(8)
    md, sd = g.eval_sources("eikonal", p)
    assert g.last_timing()["launches"][0] >= 2                 # the solver launch is booked with the discretisation stage
    assert np.array_equal(sh, sd) and sh[7] != 0 and (sh != 0).sum() == 1
    ok = sh == 0
    assert np.array_equal(mh[ok].view(np.uint32), md[ok].view(np.uint32))
    # the default: solves shared between the device and the host threads by a cost model that every batch corrects; whatever the
    # split, the results are the host path's
    g.set_eikonal_device(-1)
    for _ in range(3):
        ms_, ss_ = g.eval_sources("eikonal", p)
        assert np.array_equal(ss_, sh) and np.array_equal(mh[ok].view(np.uint32), ms_[ok].view(np.uint32))
