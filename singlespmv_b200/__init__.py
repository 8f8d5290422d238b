"""singlespmv_b200 -- B200-native SpMV engine behind the plugin surface of hir0shim/singleSpMV.

The product is libb200spmv.so (hand-written sm_100a CUDA behind the C-ABI of include/b200spmv.h);
this package is the thin host-side mirror of the reference's OptimizeProblem / SpMV interface.
"""
from ._lib import LIB_PATH, B200SpmvError, FORMATS, SYNTH          # noqa: F401
from .plugin import (DeviceCoo, OptimizeProblem, SpMat, SpMatOpt, SpMV, Vec, VecOpt,   # noqa: F401
                     device_count, host_register, host_unregister, reference_vectors)
