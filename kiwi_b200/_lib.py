"""ctypes binding of libkiwi_b200.so (the C ABI declared in include/kiwi_b200.h).

The library is built in-tree by kiwi_b200.build.  There is no fallback of any kind: if the shared
library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkiwi_b200.so")

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "kiwi_last_error": (C.c_char_p, []),
    "kiwi_version": (C.c_char_p, []),
    "kiwi_gfdb_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "kiwi_gfdb_destroy": (None, [C.c_void_p]),
    "kiwi_gfdb_save_array": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p]),
    "kiwi_gfdb_build_ahfull": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kiwi_gfdb_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kiwi_gfdb_read": (C.c_void_p, [C.c_char_p]),
    "kiwi_gfdb_read_hdf": (C.c_void_p, [C.c_char_p]),
    "kiwi_h5_read_root_dataset": (C.c_int, [C.c_char_p, C.c_char_p, c_int_p, c_int_p, c_int_p, c_ll_p, C.c_void_p, C.c_longlong, c_ll_p, c_int_p]),
    "kiwi_gfdb_interpolate": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "kiwi_gulunay": (C.c_int, [C.c_int, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p]),
    "kiwi_gfdb_meta": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_ll_p, c_ll_p]),
    "kiwi_gfdb_view": (C.c_int, [C.c_void_p, C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_ll_p), C.POINTER(c_float_p)]),
    "kiwi_create": (C.c_void_p, [C.c_int]),
    "kiwi_destroy": (None, [C.c_void_p]),
    "kiwi_set_database": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kiwi_set_local_interpolation": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_spacial_undersampling": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "kiwi_set_receivers": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_float_p, C.POINTER(C.c_char_p)]),
    "kiwi_switch_receiver": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "kiwi_set_source_location": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_double]),
    "kiwi_set_crust2x2": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kiwi_set_source_constraints": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p]),
    "kiwi_set_source_crustal_thickness_limit": (C.c_int, [C.c_void_p, C.c_float]),
    "kiwi_set_effective_dt": (C.c_int, [C.c_void_p, C.c_float]),
    "kiwi_set_ref_seismogram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, c_float_p]),
    "kiwi_set_misfit_method": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_misfit_taper": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p]),
    "kiwi_set_misfit_filter": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p]),
    "kiwi_set_synthetics_factor": (C.c_int, [C.c_void_p, C.c_float]),
    "kiwi_set_share_syntheses": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_mt_grid": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_eikonal_device": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_accumulation": (C.c_int, [C.c_void_p, C.c_int]),
    "kiwi_set_floating_shiftrange": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float]),
    "kiwi_get_nmisfits": (C.c_int, [C.c_void_p]),
    "kiwi_get_n_source_params": (C.c_int, [C.c_int]),
    "kiwi_eval_sources": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_int_p]),
    "kiwi_eval_sources_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, C.c_void_p, c_int_p]),
    "kiwi_global_misfits": (C.c_int, [C.c_int, C.c_int, c_float_p, c_float_p]),
    "kiwi_outer_misfits": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_int_p, c_double_p]),
    "kiwi_set_source_params": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "kiwi_get_misfits": (C.c_int, [C.c_void_p, c_float_p, C.c_int, c_int_p]),
    "kiwi_get_global_misfit": (C.c_int, [C.c_void_p, c_float_p]),
    "kiwi_get_peak_amplitudes": (C.c_int, [C.c_void_p, C.c_int, c_float_p, C.c_int, c_int_p]),
    "kiwi_get_arias_intensities": (C.c_int, [C.c_void_p, c_float_p, C.c_int, c_int_p]),
    "kiwi_eval_ground_motion": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_int_p]),
    "kiwi_set_source_params_mask": (C.c_int, [C.c_void_p, c_int_p, C.c_int]),
    "kiwi_set_source_subparams": (C.c_int, [C.c_void_p, c_float_p, C.c_int]),
    "kiwi_set_source_subparams_limits": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int]),
    "kiwi_get_source_subparams": (C.c_int, [C.c_void_p, c_float_p, C.c_int, c_int_p]),
    "kiwi_minimize_lm": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_float_p]),
    "kiwi_lmdif_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p, C.c_float, C.c_float, C.c_float, C.c_int,
                                     C.c_float, c_float_p, C.c_int, C.c_float, c_int_p, c_int_p]),
    "kiwi_get_floating_shifts": (C.c_int, [C.c_void_p, c_int_p, C.c_int, c_int_p]),
    "kiwi_shift_ref_seismogram": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "kiwi_get_cross_correlations": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, c_float_p, C.c_int, c_int_p, c_int_p]),
    "kiwi_autoshift_ref_seismogram": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, c_float_p, C.c_int, c_int_p]),
    "kiwi_get_distances": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, c_int_p]),
    "kiwi_get_source_crustal_thickness": (C.c_int, [C.c_void_p, c_float_p]),
    "kiwi_get_principal_axes": (C.c_int, [C.c_void_p, c_float_p, c_float_p]),
    "kiwi_get_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_float_p, C.c_int]),
    "kiwi_get_probe_spectrum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_int_p, c_float_p, C.c_int]),
    "kiwi_eikonal_fmm": (C.c_int, [C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p]),
    "kiwi_eikonal_fmm_device": (C.c_int, [C.c_int, c_int_p, c_int_p, C.POINTER(c_float_p), c_float_p, c_float_p, c_float_p, C.POINTER(c_float_p), c_float_p]),
    "kiwi_get_seismogram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_float_p, C.c_int]),
    "kiwi_discretize_source": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p, C.c_int, c_int_p, c_int_p]),
    "kiwi_get_indices": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, c_float_p, c_float_p, c_int_p, C.c_int, c_int_p]),
    "kiwi_get_spans": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "kiwi_trace_span": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p]),
    "kiwi_last_batch_bytes": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_int_p, c_ll_p]),
    "kiwi_last_timing": (C.c_int, [C.c_void_p, c_float_p, c_int_p]),
    "kiwi_host_alloc": (C.c_void_p, [C.c_size_t]),
    "kiwi_host_free": (None, [C.c_void_p]),
}


# kiwi_lm_fcn: int fcn(void* user, int ncols, int n, int m, float* xs, float* fvecs)
LM_FCN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p)


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "kiwi_b200: %s is missing. Build it with `python -m kiwi_b200.build` (needs nvcc); "
            "there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
