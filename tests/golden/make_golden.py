#!/usr/bin/env python
"""Regenerate tests/golden/*.npz.

The reference (Fortran) cannot be built in this image, so these fixtures are outputs of the CPU oracle
(oracle/, pinned on the reference's own known-answer tests) on the small shared scenario of
tests/scenario.py.  They freeze the oracle: any later edit of the restatement that changes a number
shows up as a diff against these files, and the CUDA path is held against the same numbers.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import scenario as sc  # noqa: E402
from oracle_lib import OracleEngine  # noqa: E402

COMPS = ["ned", "ar", "d", "neu", "cl", "wsd"]
EIK = np.array([0.1, 100, -200, 3500, 2e18, 40, 70, 20, 0, 0, 1500, 300, -200, 0.8, 0.4], np.float32)


def candidates():
    p = np.tile(sc.BILAT_SMALL, (7, 1))
    p[1, 5] += 15; p[2, 6] -= 20; p[3, 7] += 40; p[4, 3] += 500; p[5, 9] += 800; p[6, 4] *= 1.3
    p[6, 13] = 0.2
    return p


def build(engine_factory=OracleEngine):
    lat, lon, dep = sc.small_receivers(6)
    o = engine_factory()
    sc.setup(o, sc.small_db(), lat, lon, dep, COMPS)
    out = {}
    for name, stype, params in (("bilateral", "bilateral", sc.BILAT_SMALL), ("moment_tensor", "moment_tensor", sc.MT_SMALL), ("eikonal", "eikonal", EIK)):
        table, grid, n = o.discretize_source(stype, params)
        out["table_" + name] = table
        if name == "bilateral":
            out["grid_" + name] = np.asarray(grid[:3], np.int32)      # nx, ny, nt (source_bilat.f90:266-268)
        elif name == "eikonal":
            out["grid_" + name] = np.asarray(grid[:2], np.int32)      # nx, ny (source_eikonal.f90:311-312)
    o.set_source_params("bilateral", sc.BILAT_SMALL)
    for ir, ic in ((1, 1), (1, 3), (2, 2), (6, 1)):
        first, data = o.get_seismogram(ir, ic, 0)
        out["seis_%d_%d_first" % (ir, ic)] = np.int32(first)
        out["seis_%d_%d" % (ir, ic)] = data
    sc.set_refs_from(o, [o], [len(c) for c in COMPS])
    for norm in ("l2norm", "l1norm", "ampspec_l1norm"):
        o.set_misfit_method(norm)
        for ir in range(1, 7):
            o.set_misfit_taper(ir, [1.0, 1.6, 4.0, 5.2], [0, 1, 1, 0])
        m, st = o.eval_sources("bilateral", candidates())
        out["misfits_" + norm] = m
    return out


def build_extras(engine_factory=OracleEngine):
    """round-1 additions: cross-correlations and automatic shifts, probe exports, small getters, database interpolation"""
    lat, lon, dep = sc.small_receivers(6)
    o = engine_factory()
    is_oracle = engine_factory is OracleEngine
    sc.setup(o, sc.small_db_ng8(), lat, lon, dep, COMPS)
    out = {}
    src = OracleEngine()                       # the references are always the oracle's synthetics, late by 3 samples
    sc.setup(src, sc.small_db_ng8(), lat, lon, dep, COMPS)
    src.set_source_params("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(src, [o], [len(c) for c in COMPS], shift=3)
    for ir in range(1, 7):
        o.set_misfit_taper(ir, [1.0, 1.6, 4.0, 5.2], [0, 1, 1, 0])
    o.set_misfit_filter([0.2, 0.5, 2.0, 3.0], [0, 1, 1, 0])
    o.set_source_params("bilateral", candidates()[6])
    out["xcorr_2"] = o.get_cross_correlations(2, -0.5, 0.5)
    for which, proc in (("references", "tapered"), ("synthetics", "filtered")):
        first, data = o.get_probe(3, 1, which, proc)
        out["probe_%s_%s_first" % (which, proc)] = np.int32(first)
        out["probe_%s_%s" % (which, proc)] = data
    df, amp = o.get_probe(5, 2, "synthetics", "filtered", spectrum=True)
    out["spectrum_df"] = np.float32(df); out["spectrum_5_2"] = amp
    out["autoshift"] = np.asarray(o.autoshift_ref_seismogram(0, -0.6, 0.6), np.float32)
    o.set_misfit_method("l1norm")
    m, st = o.eval_sources("bilateral", candidates())
    out["misfits_after_autoshift"] = m
    d, a = o.get_distances()
    out["distances"] = d; out["azimuths"] = a
    if is_oracle:
        pax, tax = o.principal_axes(float(sc.BILAT_SMALL[5]), float(sc.BILAT_SMALL[6]), float(sc.BILAT_SMALL[7]))
    else:
        o.set_source_params("bilateral", sc.BILAT_SMALL)
        pax, tax = o.get_principal_axes()
    out["principal_axes"] = np.concatenate([pax, tax]).astype(np.float32)
    # Gulunay interpolation of a small database, horizontally by 2: a real trace, an interpolated one and the extrapolated last one
    from kiwi_b200 import Gfdb
    db = Gfdb.create(40, 2, 8, 0.1, 400.0, 400.0, 4000.0, 2000.0).build_ahfull(2700.0, 6000.0, 3464.0, nfflag=False)
    if is_oracle:
        from oracle_lib import gfdb_interpolate
        _meta, tr = gfdb_interpolate(db, 2, 1)
        get = lambda ix, iz, ig: tr[(ix, iz, ig)]
    else:
        g = db.interpolate(2, 1)
        s0, ln, off, dat = g.view()
        nz, ng = g.meta()["nz"], g.meta()["ng"]

        def get(ix, iz, ig):
            k = ((ix - 1) * nz + iz - 1) * ng + ig - 1
            return int(s0[k]), dat[off[k]:off[k] + ln[k]].copy()
    for ix, iz, ig in ((21, 1, 1), (22, 1, 1), (50, 2, 4), (80, 2, 8)):
        first, data = get(ix, iz, ig)
        out["interp_%d_%d_%d_first" % (ix, iz, ig)] = np.int32(first)
        out["interp_%d_%d_%d" % (ix, iz, ig)] = np.asarray(data, np.float32)
    return out


if __name__ == "__main__":
    # python tests/golden/make_golden.py [small_scenario] [round1_extras]   (default: both)
    which = [a for a in sys.argv[1:]] or ["small_scenario", "round1_extras"]
    for name, fn in (("small_scenario", build), ("round1_extras", build_extras)):
        if name in which:
            data = fn()
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
            print("wrote", os.path.join(HERE, name + ".npz"), {k: getattr(v, "shape", ()) for k, v in data.items()})
