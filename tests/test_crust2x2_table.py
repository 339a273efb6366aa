"""kiwi_b200/data/crust2x2.kcr (what the engine and the oracle load) against a direct parse of the reference's CRUST2.0 text files
as crust2x2.f90:240-341 load_crustal_model reads them.  The converter (tools/make_crust2x2_table.py) is not used here: a shared bug in it
would otherwise be invisible to parity.  Needs the reference tree (this container only); skipped where it is absent."""
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/aux/crust2x2"
KCR = os.path.join(ROOT, "kiwi_b200", "data", "crust2x2.kcr")
NLAYERS, NTYPES, NLA, NLO = 7, 360, 90, 180

pytestmark = pytest.mark.skipif(not os.path.isdir(SRC), reason="reference tree not present")


def fortran_tokens(line):
    """list-directed read: blank / comma / tab separated items"""
    return line.replace(",", " ").split()


def parse_reference_files():
    key = open(os.path.join(SRC, "CNtype2_key.txt")).read().splitlines()
    pos = 5                                    # do i=1,5: read(u,*)
    ids, vp, vs, rho, th = [], [], [], [], []
    for _ in range(NTYPES):
        ids.append(fortran_tokens(key[pos])[0][:2])                       # character(len=2) :: id
        vp.append([np.float32(t) for t in fortran_tokens(key[pos + 1])[:NLAYERS + 1]])
        vs.append([np.float32(t) for t in fortran_tokens(key[pos + 2])[:NLAYERS + 1]])
        rho.append([np.float32(t) for t in fortran_tokens(key[pos + 3])[:NLAYERS + 1]])
        th.append([np.float32(t) for t in fortran_tokens(key[pos + 4])[:NLAYERS]])
        pos += 5
    tmap = np.zeros((NLA, NLO), np.int64)
    rows = open(os.path.join(SRC, "CNtype2.txt")).read().splitlines()[1:]
    for j in range(NLA):
        toks = fortran_tokens(rows[j])[1:1 + NLO]                        # read(u,*) ilat, ctype_ids
        for i, t in enumerate(toks):
            tmap[j, i] = next(l for l in range(NTYPES) if ids[l] == t[:2])   # type_loop: first match
    elev = np.zeros((NLA, NLO), np.float32)
    rows = open(os.path.join(SRC, "CNelevatio2.txt")).read().splitlines()[1:]
    for j in range(NLA):
        elev[j] = [np.float32(t) for t in fortran_tokens(rows[j])[1:1 + NLO]]
    return np.array(vp, np.float32), np.array(vs, np.float32), np.array(rho, np.float32), np.array(th, np.float32), tmap, elev


def read_kcr():
    raw = open(KCR, "rb").read()
    assert raw[:4] == b"KCR1"
    ntypes, nlo, nla = struct.unpack("<iii", raw[4:16])
    assert (ntypes, nlo, nla) == (NTYPES, NLO, NLA)
    o = 16
    table = np.frombuffer(raw, np.float32, ntypes * 31, o).reshape(ntypes, 31); o += ntypes * 31 * 4
    tmap = np.frombuffer(raw, np.int16, nla * nlo, o).reshape(nla, nlo); o += nla * nlo * 2
    elev = np.frombuffer(raw, np.float32, nla * nlo, o).reshape(nla, nlo); o += nla * nlo * 4
    assert o == len(raw)
    return table, tmap, elev


def test_table_equals_the_reference_files():
    vp, vs, rho, th, tmap, elev = parse_reference_files()
    table, kmap, kelev = read_kcr()
    assert np.array_equal(table[:, 0:8], vp) and np.array_equal(table[:, 8:16], vs) and np.array_equal(table[:, 16:24], rho)
    assert np.array_equal(table[:, 24:31], th)
    assert np.array_equal(kmap.astype(np.int64), tmap)
    assert np.array_equal(kelev, elev)


def test_crustal_thickness_through_the_oracle_equals_a_direct_evaluation():
    """get_source_crustal_thickness (parameterized_source.f90:207-221 -> crust2x2_get_profile_averages crust2x2.f90:139-166) at a
    few hundred locations: loader (unit conversion, ice/water swap), cell lookup (latlon2indices :197-213) and the layer sum"""
    from oracle_lib import OracleEngine
    vp, vs, rho, th, tmap, elev = parse_reference_files()
    o = OracleEngine()
    rng = np.random.default_rng(4)
    for lat, lon in [(30.0, 70.0), (-89.5, 179.5), (89.9, -179.9), (0.0, 0.0)] + [(float(a), float(b)) for a, b in zip(rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300))]:
        o.set_source_location(lat, lon, 0.0)
        got = o.get_source_crustal_thickness()
        # latlon2indices on real(lat), real(lon) in degrees
        flat, flon = np.float32(lat), np.float32(lon)
        dx = np.float32(360.0) / np.float32(NLO)
        ilat = int((np.float32(90.0) - flat) / dx)
        ilon = int((flon + np.float32(180.0)) / dx)
        t = th[tmap[min(ilat, NLA - 1), min(ilon, NLO - 1)]] * np.float32(1000.0)
        t = t.copy(); t[0], t[1] = t[1], t[0]                                # flip ice and water layers
        want = np.float32(0.0)
        for i in range(1, NLAYERS):                                          # do i=2,nlayers
            want = np.float32(want + t[i])
        assert got == want, (lat, lon, got, want)
