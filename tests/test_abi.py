"""The C-ABI library loads without a GPU and exports every symbol include/kiwi_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kiwi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kiwi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kiwi_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the Python binding table covers the header one to one
    assert sorted(_lib.SIGNATURES) == names


def test_fortran_binding_covers_the_engine_facing_entry_points():
    """fortran/kiwi_b200_binding.f90 (untested source, no Fortran compiler here) declares an interface for every entry point a Kiwi
    maintainer would call from minimizer_engine.f90; what it leaves out are the inspection, measurement and tool entry points"""
    names = set(declared_symbols())
    src = open(os.path.join(ROOT, "fortran", "kiwi_b200_binding.f90")).read()
    bound = set(re.findall(r'name="(kiwi_[a-z0-9_]+)"', src))
    assert bound <= names, sorted(bound - names)
    tools = {"kiwi_discretize_source", "kiwi_eikonal_fmm", "kiwi_eikonal_fmm_device", "kiwi_eval_sources_device", "kiwi_get_indices", "kiwi_get_n_source_params", "kiwi_get_spans",
             "kiwi_gfdb_build_ahfull", "kiwi_gfdb_meta", "kiwi_gfdb_view", "kiwi_gfdb_write", "kiwi_global_misfits", "kiwi_gulunay",
             "kiwi_h5_read_root_dataset", "kiwi_host_alloc", "kiwi_host_free", "kiwi_last_batch_bytes", "kiwi_last_timing", "kiwi_lmdif_batched",
             "kiwi_trace_span", "kiwi_version"}
    assert names - bound == tools, sorted((names - bound) ^ tools)
    assert src.count("end function") + src.count("end subroutine") == len(bound)


def test_version_and_parameter_counts():
    from kiwi_b200 import _lib, n_source_params
    assert b"sm_100a" in _lib.lib.kiwi_version()
    assert n_source_params("bilateral") == 14          # source_bilat.f90:32
    assert n_source_params("moment_tensor") == 11      # source_moment_tensor.f90
    assert n_source_params("circular") == 11           # source_circular.f90:33
    assert n_source_params("point_lp") == 13           # source_point_lp.f90:14
    assert n_source_params("eikonal") == 15 and n_source_params("mt_eikonal") == 20


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kiwi_b200 import Engine, KiwiError
    with pytest.raises(KiwiError, match="no CUDA device"):
        Engine(0)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under kiwi_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kiwi_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"oracle_lib|liboracle|oracle/|ko_[a-z]+\.hpp", text):
                    bad.append(f)
    assert not bad, bad
