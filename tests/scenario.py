"""Shared small scenarios for the parity tests: the same calls are applied to the CUDA engine
(kiwi_b200.Engine) and to the oracle (tests/oracle_lib.OracleEngine)."""
import functools

import numpy as np

from kiwi_b200 import Gfdb
from kiwi_b200 import synthetic

ORIGIN = (30.0, 70.0)


@functools.lru_cache(maxsize=None)
def small_db():
    # 0.25-20 km x 0-5.75 km, near+far field, the kiwibench source time function
    return Gfdb.create(80, 24, 10, 0.1, 250.0, 250.0, 250.0, 0.0).build_ahfull(2700.0, 6000.0, 3464.0)


@functools.lru_cache(maxsize=None)
def small_db_ng8():
    return Gfdb.create(60, 16, 8, 0.1, 300.0, 300.0, 300.0, 0.0).build_ahfull(2700.0, 6000.0, 3464.0, nfflag=False)


BILAT_SMALL = np.array([0.3, 200, -300, 3000, 1.5e18, 75, 70, 150, 20, 3000, 2000, 2000, 3000, 0.5], dtype=np.float32)
MT_SMALL = np.array([0.2, 150, -250, 2800, 1e18, -0.4e18, -0.6e18, 0.3e18, 0.2e18, -0.5e18, 0.7], dtype=np.float32)


def small_receivers(n=6, seed=7, dmin=8e3, dmax=14e3):
    lat, lon, dep = synthetic.receivers(n, ORIGIN, dmin, dmax, seed)
    return lat, lon, dep


def setup(eng, db, lat, lon, dep, comps, interpolation="bilinear", effective_dt=0.2, under=(1, 1)):
    eng.set_database(db)
    eng.set_local_interpolation(interpolation)
    eng.set_spacial_undersampling(*under)
    eng.set_receivers(lat, lon, dep, comps)
    eng.set_source_location(ORIGIN[0], ORIGIN[1], 0.0)
    eng.set_effective_dt(effective_dt)


def set_refs_from(eng_src, engines, ncomps, scale=1.07, shift=0, dt=0.1):
    """Reference traces = synthetics of the source currently held by eng_src (what
    set_synthetic_reference does, python/tunguska/seismosizer.py:523-527) scaled by `scale`."""
    refs = {}
    for ir, nc in enumerate(ncomps, start=1):
        for ic in range(1, nc + 1):
            first, data = eng_src.get_seismogram(ir, ic, 1)
            # sample i of a strip sits at index i; set_ref_seismogram places sample 0 at nint(tbegin/dt)+1
            refs[(ir, ic)] = (first, data * np.float32(scale))
    for e in engines:
        for (ir, ic), (first, data) in refs.items():
            e.set_ref_seismogram(ir, ic, (first - 1 + shift) * dt, data)
    return refs
