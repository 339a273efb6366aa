"""The `minimizer`-compatible command front-end (kiwi_b200/kiwi_minimizer): reply format of
minimizer.f90:1676-1701 and, on a GPU, a whole session held against the Python binding."""
import os
import subprocess

import numpy as np
import pytest

import scenario as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kiwi_b200", "kiwi_minimizer")


def talk(lines, env=None):
    r = subprocess.run([BIN], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    return r.stdout.splitlines()


def test_reply_format_and_line_cleaning():
    out = talk(["", "   # only a comment", "bogus_command 1 2   # trailing comment", "   frobnicate"])
    # blank and comment-only lines produce no reply (minimizer.f90:1721-1725); unknown commands answer nok + message
    assert out == ["bogus_command: nok >", "unknown command: bogus_command", "frobnicate: nok >", "unknown command: frobnicate"]


def test_no_cpu_path_behind_the_protocol():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = talk(["set_effective_dt 0.5"])
    assert out[0] == "set_effective_dt: nok >" and "no CUDA device" in out[1]


@pytest.mark.gpu
def test_session_matches_python_binding(tmp_path):
    from kiwi_b200 import Engine
    db = sc.small_db()
    dbfile = tmp_path / "db.kgf1"
    db.write(dbfile)
    lat, lon, dep = sc.small_receivers(4)
    comps = ["ned", "d", "ar", "neu"]
    rfile = tmp_path / "receivers.table"
    with open(rfile, "w") as f:
        f.write("# lat lon depth components\n")
        for a, b, c, d in zip(lat, lon, dep, comps):
            f.write("%.10f %.10f %g %s\n" % (a, b, c, d))
    p = " ".join("%.9g" % v for v in sc.BILAT_SMALL)
    p2 = sc.BILAT_SMALL.copy(); p2[5] += 20
    base = str(tmp_path / "ref")
    cands = tmp_path / "cands.txt"
    np.savetxt(cands, np.stack([sc.BILAT_SMALL, p2]), fmt="%.9g")
    env = dict(os.environ, KIWI_CRUST2X2=os.path.join(ROOT, "kiwi_b200", "data", "crust2x2.kcr"))
    out = talk(["set_database %s" % dbfile, "set_local_interpolation bilinear", "set_receivers %s has_depth" % rfile,
                "set_source_location %g %g 0" % sc.ORIGIN, "set_effective_dt 0.2", "set_source_params bilateral " + p,
                "get_misfits",                                   # no references yet
                "output_seismograms %s table synthetics plain" % base, "set_ref_seismograms %s table" % base,
                "set_misfit_method l1norm", "set_misfit_taper 1 1.0 0 1.6 1 4.0 1 5.2 0",
                "set_source_params bilateral " + " ".join("%.9g" % v for v in p2), "get_misfits", "get_global_misfit",
                "eval_sources bilateral %s" % cands, "switch_receiver 2 off", "get_misfits", "set_misfit_method nonsense",
                "output_source_model %s" % (tmp_path / "model")], env=env)
    it = iter(out)
    for _ in range(6):
        assert next(it).endswith(": ok")
    assert next(it) == "get_misfits: nok >" and next(it) == "no reference seismograms set"
    for _ in range(5):
        assert next(it).endswith(": ok")
    assert next(it) == "get_misfits: ok >"
    mis = np.array(next(it).split(), np.float32).reshape(-1, 2)
    assert next(it) == "get_global_misfit: ok >"
    gm = float(next(it))
    assert next(it) == "eval_sources: ok >"
    gms = np.array(next(it).split(), np.float32)
    assert next(it) == "switch_receiver: ok"
    assert next(it) == "get_misfits: ok >"
    mis_off = np.array(next(it).split(), np.float32).reshape(-1, 2)
    assert next(it) == "set_misfit_method: nok >" and next(it) == "unknown norm method: nonsense"
    assert next(it) == "output_source_model: ok"
    # the same through the Python binding: references = the table files the front-end wrote (text round trip)
    e = Engine(0)
    sc.setup(e, db, lat, lon, dep, comps)
    for ir, c in enumerate(comps, 1):
        for ic, ch in enumerate(c, 1):
            t = np.loadtxt("%s-%d-%s.table" % (base, ir, ch))
            e.set_ref_seismogram(ir, ic, np.float32(t[0, 0]), t[:, 1].astype(np.float32))
    e.set_misfit_method("l1norm"); e.set_misfit_taper(1, [1.0, 1.6, 4.0, 5.2], [0, 1, 1, 0])
    e.set_source_params("bilateral", p2)
    want = e.get_misfits()
    assert mis.shape == want.shape == (9, 2)
    assert np.allclose(mis, want, rtol=1e-6, atol=0)
    assert abs(gm - e.get_global_misfit()) <= 1e-6 * gm
    assert abs(gms[1] - gm) <= 1e-6 * gm and gms[0] < 1e-3 * gm       # candidate 0 is the reference itself
    assert mis_off.shape == (8, 2)
    # output_source_model (minimizer_engine.f90:948-978): the centroid table of the current source and its size
    table, grid, n = e.discretize_source("bilateral", p2)
    dsm = np.loadtxt(str(tmp_path / "model") + "-dsm.table", dtype=np.float32)
    assert dsm.shape == (n, 10) and np.array_equal(dsm, table)
    info = open(str(tmp_path / "model") + "-tdsm.info").read().split()
    assert info == ["ncentroids", str(n)]
    assert open(str(tmp_path / "model") + "-psm.info").read().split()[0] == "origin"


@pytest.mark.gpu
def test_session_with_interpolated_database_and_autoshift(tmp_path):
    """`set_database dbpath nipx nipz`, `shift_ref_seismogram`, `autoshift_ref_seismogram` (minimizer.f90:89-135, 356-386, 447-483)"""
    from kiwi_b200 import Engine
    db = sc.small_db_ng8()
    dbfile = tmp_path / "db.kgf1"
    db.write(dbfile)
    lat, lon, dep = sc.small_receivers(3)
    rfile = tmp_path / "receivers.table"
    with open(rfile, "w") as f:
        for a, b, c in zip(lat, lon, dep):
            f.write("%.10f %.10f %g ned\n" % (a, b, c))
    p = " ".join("%.9g" % v for v in sc.BILAT_SMALL)
    base = str(tmp_path / "ref")
    out = talk(["set_database %s 2 1" % dbfile, "set_local_interpolation bilinear", "set_receivers %s has_depth" % rfile,
                "set_source_location %g %g 0" % sc.ORIGIN, "set_effective_dt 0.2", "set_source_params bilateral " + p,
                "output_seismograms %s table synthetics plain" % base, "set_ref_seismograms %s table" % base,
                "shift_ref_seismogram 2 0.3", "get_global_misfit", "autoshift_ref_seismogram 0 -0.5 0.5", "get_global_misfit",
                "autoshift_ref_seismogram 9 -0.5 0.5", "shift_ref_seismogram 1",
                "output_cross_correlations %s -0.2 0.2" % (base + "cc"), "get_cached_traces_memory", "set_verbose T",
                "set_misfit_filter_1 2 0.2 0 0.5 1 2.0 1 3.0 0", "get_global_misfit", "get_principal_axes",
                "output_distances %s" % (base + ".distances"), "set_misfit_taper 1 1.0 0 1.6 1 4.0 1 5.2 0",
                "output_seismograms %s table references tapered" % (base + "rt"), "output_seismogram_spectra %s synthetics filtered" % (base + "sp"),
                "output_seismograms %s table nonsense plain" % base])
    it = iter(out)
    for _ in range(9):
        assert next(it).endswith(": ok"), out
    assert next(it) == "get_global_misfit: ok >"
    gm_shifted = float(next(it))
    assert next(it) == "autoshift_ref_seismogram: ok >"
    shifts = np.array(next(it).split(), np.float32)
    assert next(it) == "get_global_misfit: ok >"
    gm_back = float(next(it))
    assert next(it) == "autoshift_ref_seismogram: nok >" and next(it) == "receiver index out of range"
    assert next(it) == "shift_ref_seismogram: nok >" and next(it).startswith("usage: shift_ref_seismogram")
    assert next(it) == "output_cross_correlations: ok"
    assert next(it) == "get_cached_traces_memory: ok >" and int(next(it)) > 4 * db.meta()["nsamples"] * 1.9      # interpolated: twice the traces
    assert next(it) == "set_verbose: ok" and next(it) == "set_misfit_filter_1: ok"
    assert next(it) == "get_global_misfit: ok >" and float(next(it)) < 1e-3
    assert next(it) == "get_principal_axes: ok >" and len(next(it).split()) == 4
    assert next(it) == "output_distances: ok"
    dist = np.loadtxt(base + ".distances")
    assert dist.shape == (3, 3) and np.all((dist[:, 1] > 7e3) & (dist[:, 1] < 15e3)) and np.allclose(dist[:, 0], dist[:, 1] / 6371e3 * 180 / np.pi, rtol=1e-6)
    assert next(it) == "set_misfit_taper: ok" and next(it) == "output_seismograms: ok" and next(it) == "output_seismogram_spectra: ok"
    assert next(it) == "output_seismograms: nok >" and next(it) == "unknown probe name: nonsense"
    rt, plain = np.loadtxt(base + "rt-1-n.table"), np.loadtxt(base + "-1-n.table")
    assert rt[0, 0] >= 1.0 - 1e-6 and rt[-1, 0] <= 5.2 + 1e-6 and rt.shape[0] < plain.shape[0] and np.abs(rt[:, 1]).max() <= np.abs(plain[:, 1]).max() * 1.07 * (1 + 1e-6)
    sp = np.loadtxt(base + "sp-2-e.table")
    assert sp[0, 0] == 0.0 and np.all(np.diff(sp[:, 0]) > 0) and sp[-1, 0] == pytest.approx(5.0) and sp[np.argmax(sp[:, 1]), 0] < 3.0
    cc = np.loadtxt(base + "cc-2-e.table")
    assert cc.shape == (5, 2) and np.allclose(cc[:, 0], [-0.2, -0.1, 0.0, 0.1, 0.2], atol=1e-6) and np.argmax(cc[:, 1]) == 2
    assert np.allclose(shifts, [0.0, -0.3, 0.0], atol=1e-6)          # the far-field traces correlate best where they came from
    assert gm_shifted > 0.1 and gm_back < 1e-4
    # the interpolated database is the one the binding builds
    e = Engine(0)
    e.set_database(db, 2, 1)
    m = e._db.meta()
    assert m["nx"] == 2 * db.meta()["nx"] and abs(m["dx"] - db.meta()["dx"] / 2) < 1e-3


@pytest.mark.gpu
def test_set_accumulation_command(tmp_path):
    """the front-end's extension `set_accumulation reference|batched` (kiwi_set_accumulation): with `reference` the seismogram files hold
    the oracle's samples bit for bit"""
    from oracle_lib import OracleEngine
    db = sc.small_db()
    dbfile = tmp_path / "db.kgf1"
    db.write(dbfile)
    lat, lon, dep = sc.small_receivers(2)
    comps = ["ned", "ar"]
    rfile = tmp_path / "receivers.table"
    with open(rfile, "w") as f:
        for a, b, c, d in zip(lat, lon, dep, comps):
            f.write("%.10f %.10f %g %s\n" % (a, b, c, d))
    p = " ".join("%.9g" % v for v in sc.BILAT_SMALL)
    base = str(tmp_path / "syn")
    out = talk(["set_database %s" % dbfile, "set_local_interpolation bilinear", "set_receivers %s has_depth" % rfile,
                "set_source_location %g %g 0" % sc.ORIGIN, "set_effective_dt 0.2", "set_accumulation nonsense", "set_accumulation reference",
                "set_source_params bilateral " + p, "output_seismograms %s table synthetics plain" % base, "set_accumulation batched"])
    assert "set_accumulation: nok >" in out and out.count("set_accumulation: ok") == 2 and "output_seismograms: ok" in out, out
    o = OracleEngine()
    sc.setup(o, db, lat, lon, dep, comps)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    first, data = o.get_seismogram(1, 1, 1)
    tab = np.loadtxt(base + "-1-n.table")
    assert tab.shape[0] == data.size and np.array_equal(tab[:, 1].astype(np.float32), data)      # (%.9g: every fp32 value survives the file)
