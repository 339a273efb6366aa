"""Kiwi's HDF5 Green's function database read without libhdf5: kiwi_gfdb_read_hdf against files
written by the independent minimal writer tests/h5mini_writer.py (same published format specification; no real HDF5
library exists in this image, so these files are the only fixtures: the reader is NOT pinned on library-written files)."""
import os

import numpy as np
import pytest

from kiwi_b200 import Gfdb, KiwiError
import h5mini_writer as h5w


def random_traces(rng, nx, nz, ng, p_missing=0.15):
    traces, dense = {}, {}
    for ix in range(1, nx + 1):
        for iz in range(1, nz + 1):
            for ig in range(1, ng + 1):
                if rng.random() < p_missing:
                    continue
                nstrips = int(rng.integers(1, 4))
                first = int(rng.integers(1, 60))
                strips, pos = [], first
                for s in range(nstrips):
                    n = int(rng.integers(3, 40))
                    d = rng.standard_normal(n).astype(np.float32)
                    d[0] = d[0] or 1.0; d[-1] = d[-1] or 1.0
                    strips.append((pos, d))
                    pos += n + int(rng.integers(1, 12))            # zeros between the strips (sparse_trace.f90:29-50)
                last = strips[-1][0] + strips[-1][1].size
                full = np.zeros(last - first, np.float32)
                for (o, d) in strips:
                    full[o - first:o - first + d.size] = d
                traces[(ix, iz, ig)] = strips
                dense[(ix, iz, ig)] = (first, full)
    return traces, dense


@pytest.mark.parametrize("nx,nz,ng,nxc,with_first", [(5, 3, 10, 2, True), (3, 2, 8, 3, False), (40, 12, 10, 16, True)])
def test_read_kiwi_hdf_database(tmp_path, nx, nz, ng, nxc, with_first):
    rng = np.random.default_rng(nx * 100 + nz)
    traces, dense = random_traces(rng, nx, nz, ng)
    base = str(tmp_path / "db")
    h5w.write_kiwi_gfdb(base, nx, nz, ng, 0.5, 2000.0, 1000.0, 4000.0, 500.0, nxc, traces, with_first)
    nchunks = -(-nx // nxc)
    assert os.path.exists(base + ".index") and all(os.path.exists("%s.%d.chunk" % (base, i)) for i in range(1, nchunks + 1))
    got = Gfdb.read_hdf(base)
    want = Gfdb.create(nx, nz, ng, 0.5, 2000.0, 1000.0, 4000.0 if with_first else 0.0, 500.0 if with_first else 0.0)
    for (ix, iz, ig), (first, full) in dense.items():
        want.save_array(ix, iz, ig, first, full)
    mg, mw = got.meta(), want.meta()
    assert mg == mw
    (s0g, lg, og, dg), (s0w, lw, ow, dw) = got.view(), want.view()
    assert np.array_equal(s0g, s0w) and np.array_equal(lg, lw) and np.array_equal(og, ow)
    assert np.array_equal(dg.view(np.uint32), dw.view(np.uint32))          # samples bit for bit
    assert mg["ntraces"] == len(dense)


def test_errors_are_reported(tmp_path):
    base = str(tmp_path / "nodb")
    with pytest.raises(KiwiError, match="failed to open file"):
        Gfdb.read_hdf(base)
    rng = np.random.default_rng(1)
    traces, _ = random_traces(rng, 2, 2, 10)
    base = str(tmp_path / "db")
    h5w.write_kiwi_gfdb(base, 2, 2, 10, 0.5, 1.0, 1.0, 0.0, 0.0, 2, traces)
    raw = open(base + ".1.chunk", "rb").read()
    open(base + ".1.chunk", "wb").write(raw[:len(raw) // 2])            # truncated chunk file
    with pytest.raises(KiwiError, match="outside of file|signature|not a dataset"):
        Gfdb.read_hdf(base)
    bad = bytearray(raw); bad[8] = 2                                     # superblock version 2 (HDF5 1.10 'latest' format)
    open(base + ".1.chunk", "wb").write(bytes(bad))
    with pytest.raises(KiwiError, match="superblock version 2 is not supported"):
        Gfdb.read_hdf(base)
    os.remove(base + ".1.chunk")
    with pytest.raises(KiwiError, match="failed to open file"):
        Gfdb.read_hdf(base)
    open(base + ".index", "wb").write(b"not an hdf5 file" * 8)
    with pytest.raises(KiwiError, match="not an HDF5 file"):
        Gfdb.read_hdf(base)


def test_parser_on_a_file_written_by_the_real_hdf5_library():
    """The only HDF5 file in this image that libhdf5 itself wrote: scipy's MATLAB 7.3 test file (MATLAB 7.4, 2008, HDF5 1.6:
    superblock 0 behind a 512-byte user block, symbol-table root group, version-1 object header, one double dataset with
    a MATLAB_class attribute).  It holds testdouble = 0:pi/4:2*pi as a 9 x 1 array."""
    import scipy.io.matlab
    from kiwi_b200 import h5_read_root_dataset, h5_root_members
    path = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    assert h5_root_members(path) == ["testdouble"]
    a, nattrs = h5_read_root_dataset(path, "testdouble")
    assert a.dtype == np.float64 and a.shape == (9, 1) and nattrs == 1
    assert np.allclose(a[:, 0], np.arange(9) * np.pi / 4, rtol=1e-15, atol=0)
    with pytest.raises(KiwiError, match="no object 'nothing'"):
        h5_read_root_dataset(path, "nothing")
