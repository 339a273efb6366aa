"""kiwi_b200 -- B200-native forward-modelling and misfit engine for Kiwi's source-inversion loop.

Host-side mirror of the reference's command surface (minimizer.f90 / minimizer_engine.f90) on top
of the C ABI in include/kiwi_b200.h.  All computation happens in the hand-written sm_100a kernels
of libkiwi_b200.so; this package only marshals arguments.

The names below resolve lazily so that `python -m kiwi_b200.build` can run before the shared
library exists; anything else fails loudly if the library is missing (kiwi_b200/_lib.py).
"""
__all__ = ["Engine", "Gfdb", "KiwiError", "SOURCE_TYPES", "NORMS", "KIWIBENCH_STF", "n_source_params",
           "global_misfits", "lmdif_batched", "h5_root_members", "h5_read_root_dataset", "MisfitGrid"]


def __getattr__(name):
    if name == "MisfitGrid":
        from .grid_search import MisfitGrid
        return MisfitGrid
    if name in __all__:
        from . import engine
        return getattr(engine, name)
    raise AttributeError("module 'kiwi_b200' has no attribute %r" % name)
