"""Host-side database container, KGF1 file and the analytical full-space builder (fixture generator)."""
import numpy as np
import pytest

import scenario as sc
from kiwi_b200 import Gfdb, KiwiError
from oracle_lib import OracleEngine


def test_save_array_packs_like_trace_pack():
    # sparse_trace.f90:443-555: leading zeros dropped, exactly one trailing zero kept
    db = Gfdb.create(2, 2, 10, 0.5, 100, 100, 100, 0)
    db.save_array(1, 1, 1, 20, [0, 0, 0, 1, 2, 3, 0, 0, 0])
    db.save_array(1, 1, 2, 5, [0, 0, 0, 0])
    db.save_array(2, 2, 10, -3, [4, 5])
    span0, length, offset, data = db.view()
    assert (span0[0], length[0]) == (23, 4)
    assert list(data[offset[0]:offset[0] + 4]) == [1, 2, 3, 0]
    assert (span0[1], length[1]) == (5, 1) and data[offset[1]] == 0
    i = ((2 - 1) * 2 + (2 - 1)) * 10 + 9
    assert (span0[i], length[i]) == (-3, 2)
    assert db.meta()["ntraces"] == 3
    with pytest.raises(KiwiError, match="out of bounds"):
        db.save_array(3, 1, 1, 0, [1.0])


def test_spans_agree_with_oracle_trace_pack():
    db = sc.small_db()
    o = OracleEngine()
    o.set_database(db)
    span0, length, _, _ = db.view()
    m = db.meta()
    rng = np.random.default_rng(0)
    for _ in range(200):
        ix, iz, ig = int(rng.integers(1, m["nx"] + 1)), int(rng.integers(1, m["nz"] + 1)), int(rng.integers(1, 11))
        i = ((ix - 1) * m["nz"] + (iz - 1)) * 10 + (ig - 1)
        s, _n = o.trace_span(ix, iz, ig)
        assert (s[0], s[1]) == (span0[i], span0[i] + length[i] - 1)


def test_kgf1_round_trip(tmp_path):
    db = sc.small_db_ng8()
    path = tmp_path / "db.kgf1"
    db.write(path)
    db2 = Gfdb.read(path)
    assert db.meta() == db2.meta()
    for a, b in zip(db.view(), db2.view()):
        assert np.array_equal(a, b)
    with pytest.raises(KiwiError):
        Gfdb.read(tmp_path / "missing.kgf1")


def test_ahfull_arrival_times_and_far_field_amplitude():
    """Physics check of the fixture generator (elseis.f90:133-209): P and S onsets at d/alpha, d/beta,
    far-field P amplitude of the vertical-dip-slip-free component follows 1/(4 pi rho alpha^3 r)."""
    rho, alpha, beta, dt = 2700.0, 6000.0, 3464.0, 0.1
    db = Gfdb.create(40, 4, 10, dt, 500.0, 500.0, 500.0, 0.0).build_ahfull(rho, alpha, beta, nfflag=False)
    span0, length, offset, data = db.view()
    m = db.meta()
    for ix in (10, 25, 40):
        iz = 1
        x = 500.0 * ix
        i = ((ix - 1) * m["nz"] + (iz - 1)) * 10 + 0   # g1: radial, source a (mxx=myy-like term)
        assert length[i] > 1
        t0 = span0[i] * dt
        # the stored trace starts where the far-field P pulse becomes non-zero: P travel time plus the
        # 5 leading samples over which the kiwibench STF (and its central-difference derivative) is zero
        assert abs(t0 - (x / alpha + 5 * dt)) <= 2 * dt
        tr = data[offset[i]:offset[i] + length[i]]
        # far-field P displacement for a ramp STF reaching 1: plateau of d(stf)/dt = 1/(10 dt) per sample
        nP = int(round((x / beta - x / alpha) / dt))
        peakP = np.abs(tr[:max(nP - 2, 3)]).max()
        expect = 1.0 / (4 * np.pi * rho * alpha ** 3 * x) * (1.0 / (10 * dt))
        assert 0.5 * expect < peakP < 2.0 * expect


def test_near_field_static_offset_is_stored_as_last_sample():
    db = sc.small_db()
    span0, length, offset, data = db.view()
    m = db.meta()
    i = ((8 - 1) * m["nz"] + (6 - 1)) * 10 + 0
    tr = data[offset[i]:offset[i] + length[i]]
    assert tr[-1] != 0.0 and abs(tr[-1] - tr[-2]) < 1e-3 * abs(tr[-1])   # static displacement reached


def test_reference_round_trip_known_answer(tmp_path):
    """test_gfdb.f90:33-80: a 2-chunk 3 x 2 x 8 database, one trace of two strips [11,17] = 1 0 0 0 1 1 1 and [23,24] = 1 1 saved at
    (1, 2, 2) and recovered -- here through the host container, the KGF1 file and (via the independent writer) Kiwi's HDF5 layout;
    the container holds the trace densely over its span with the gap between the strips as zeros"""
    import h5mini_writer as h5w
    s1, s2 = np.array([1, 0, 0, 0, 1, 1, 1], np.float32), np.array([1, 1], np.float32)
    dense = np.zeros(24 - 11 + 1, np.float32)
    dense[0:7] = s1; dense[23 - 11:] = s2
    db = Gfdb.create(3, 2, 8, 1.0, 1.0, 1.0, 0.0, 0.0)
    db.save_array(1, 2, 2, 11, dense)
    path = tmp_path / "test-db.kgf1"
    db.write(path)
    h5w.write_kiwi_gfdb(str(tmp_path / "test-db"), 3, 2, 8, 1.0, 1.0, 1.0, 0.0, 0.0, 2, {(1, 2, 2): [(11, s1), (23, s2)]})     # nchunks = 2
    for got in (db, Gfdb.read(path), Gfdb.read_hdf(str(tmp_path / "test-db"))):
        m = got.meta()
        assert (m["nx"], m["nz"], m["ng"]) == (3, 2, 8) and m["ntraces"] == 1            # "reopen index"
        span0, length, offset, data = got.view()
        i = ((1 - 1) * 2 + (2 - 1)) * 8 + (2 - 1)
        assert (span0[i], length[i]) == (11, 14)
        assert np.array_equal(data[offset[i]:offset[i] + 14], dense)                      # "recover 2" .. "recover 5"
        assert length[((2 - 1) * 2 + (2 - 1)) * 8 + 1] == 0                               # a trace that was never saved stays absent
    o = OracleEngine()
    o.set_database(db)
    (a, b), nstrips = o.trace_span(1, 2, 2)
    assert (a, b, nstrips) == (11, 24, 1)      # same span; one strip: a gap of 5 zeros does not exceed maxgap (sparse_trace.f90:24, :466-470)


def test_kgf1_reader_refuses_corrupt_files(tmp_path):
    """a truncated or tampered dump is an error message, not a crash or an out-of-bounds read later in set_database"""
    import struct
    db = sc.small_db_ng8()
    path = tmp_path / "db.kgf1"
    db.write(path)
    raw = path.read_bytes()
    m = db.meta()
    ntr = m["nx"] * m["nz"] * m["ng"]

    def check(blob, what):
        bad = tmp_path / "bad.kgf1"
        bad.write_bytes(blob)
        with pytest.raises(KiwiError, match=what):
            Gfdb.read(bad)
    check(raw[:len(raw) // 2], "file size")                                        # truncated
    check(raw + b"\0" * 16, "file size")                                           # trailing garbage
    check(raw[:8] + struct.pack("<i", 1 << 30) + raw[12:], "invalid header|file size")   # nx huge
    check(raw[:8] + struct.pack("<i", -3) + raw[12:], "invalid header")            # nx negative
    check(raw[:16] + struct.pack("<i", 9) + raw[20:], "invalid header")            # ng = 9
    check(raw[:24] + struct.pack("<f", 0.0) + raw[28:], "invalid header")          # dt = 0
    hdr = 56
    off_table = hdr + 2 * 4 * ntr
    check(raw[:off_table] + struct.pack("<q", 1 << 40) + raw[off_table + 8:], "trace table")     # offset beyond the samples
    len_table = hdr + 4 * ntr
    check(raw[:len_table] + struct.pack("<i", -5) + raw[len_table + 4:], "trace table")          # negative length
    check(b"JUNK" + raw[4:], "not a KGF1")


def test_ahfull_builder_against_the_independent_restatement():
    """the product's fixture generator (csrc/gfdb_host.cpp) and the oracle's (oracle/ko_ahfull.hpp) were written separately from
    gfdb_build_ahfull.f90 / elseis.f90; every trace of a node must come out bit for bit the same (span, length, samples) -- the
    database is what BOTH the CUDA path and the oracle consume in the parity tests"""
    from kiwi_b200.engine import KIWIBENCH_STF
    from oracle_lib import ahfull_node
    rng = np.random.default_rng(2)
    for db, (rho, alpha, beta), nf in ((sc.small_db(), (2700.0, 6000.0, 3464.0), True), (sc.small_db_ng8(), (2700.0, 6000.0, 3464.0), False)):
        m = db.meta()
        span0, length, offset, data = db.view()
        nodes = {(1, 1), (m["nx"], m["nz"]), (1, m["nz"]), (m["nx"], 1)} | {(int(rng.integers(1, m["nx"] + 1)), int(rng.integers(1, m["nz"] + 1))) for _ in range(40)}
        for ix, iz in sorted(nodes):
            x = np.float32(m["firstx"]) + np.float32(ix - 1) * np.float32(m["dx"])
            z = np.float32(m["firstz"]) + np.float32(iz - 1) * np.float32(m["dz"])
            want = ahfull_node(rho, alpha, beta, KIWIBENCH_STF, m["dt"], float(x), float(z), nf, True)
            for ig in range(1, m["ng"] + 1):
                i = ((ix - 1) * m["nz"] + (iz - 1)) * m["ng"] + (ig - 1)
                s0, d = want[ig - 1]
                assert (span0[i], length[i]) == (s0, d.size), (ix, iz, ig)
                assert np.array_equal(data[offset[i]:offset[i] + length[i]].view(np.uint32), d.view(np.uint32)), (ix, iz, ig)


def test_ahfull_builder_bench_l_nodes_against_the_restatement():
    """the same on nodes of the benchmark database's geometry (bench-L: dx 100 m, dz 200 m, dt 0.1 s)"""
    from kiwi_b200.engine import KIWIBENCH_STF
    from oracle_lib import ahfull_node
    db = Gfdb.create(40, 12, 10, 0.1, 100.0, 200.0, 100.0 + 1500 * 100.0, 0.0).build_ahfull(2700.0, 6000.0, 3464.0, KIWIBENCH_STF)
    m = db.meta()
    span0, length, offset, data = db.view()
    for ix in (1, 17, 40):
        for iz in (1, 6, 12):
            x = np.float32(m["firstx"]) + np.float32(ix - 1) * np.float32(m["dx"])
            z = np.float32(m["firstz"]) + np.float32(iz - 1) * np.float32(m["dz"])
            want = ahfull_node(2700.0, 6000.0, 3464.0, KIWIBENCH_STF, 0.1, float(x), float(z), True, True)
            for ig in range(1, 11):
                i = ((ix - 1) * m["nz"] + (iz - 1)) * 10 + (ig - 1)
                s0, d = want[ig - 1]
                assert (span0[i], length[i]) == (s0, d.size), (ix, iz, ig)
                assert np.array_equal(data[offset[i]:offset[i] + length[i]].view(np.uint32), d.view(np.uint32)), (ix, iz, ig)
