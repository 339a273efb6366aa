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


if __name__ == "__main__":
    data = build()
    np.savez_compressed(os.path.join(HERE, "small_scenario.npz"), **data)
    print("wrote", os.path.join(HERE, "small_scenario.npz"), {k: getattr(v, "shape", ()) for k, v in data.items()})
