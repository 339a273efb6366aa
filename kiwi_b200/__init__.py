"""kiwi_b200 -- B200-native forward-modelling and misfit engine for Kiwi's source-inversion loop.

Host-side mirror of the reference's command surface (minimizer.f90 / minimizer_engine.f90) on top
of the C ABI in include/kiwi_b200.h.  All computation happens in the hand-written sm_100a kernels
of libkiwi_b200.so; this package only marshals arguments.
"""
from .engine import (Engine, Gfdb, KiwiError, SOURCE_TYPES, NORMS, KIWIBENCH_STF, n_source_params,
                     global_misfits)

__all__ = ["Engine", "Gfdb", "KiwiError", "SOURCE_TYPES", "NORMS", "KIWIBENCH_STF", "n_source_params",
           "global_misfits"]
