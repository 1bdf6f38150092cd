"""B200-native batched state validity / planning core for the NAO whole-body RRT-Connect planner.

The product is libnwb.so (C++ host code + hand-written sm_100a CUDA kernels behind the C ABI in
include/nwb.h).  This package is the thin Python handle used by tests and benchmarks; importing it
does not load the library, creating a Context does - and fails loudly if it is not built.
"""
from ._abi import (ADVANCED, F32, F64, NFRAMES, NJ, REACHED, REASON_COLLISION, REASON_UNSTABLE, TRAPPED, NwbError,
                   NwbParams, default_params, load_library)
from .context import Context, PinnedArray, Tree, load_ds_database, write_configs

__all__ = ["Context", "PinnedArray", "Tree", "load_ds_database", "write_configs", "NwbParams", "NwbError", "default_params", "load_library", "NJ", "NFRAMES",
           "ADVANCED", "REACHED", "TRAPPED", "REASON_COLLISION", "REASON_UNSTABLE", "F64", "F32"]
