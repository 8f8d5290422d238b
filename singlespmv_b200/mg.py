"""One process driving several GPUs: ctypes binding of b200spmv_mg_* (include/b200spmv.h, csrc/mg.cu).

The reference's driver calls ONE SpMV per loop iteration (src/main.cpp:58-102); with this handle that call fans out over
the GPUs of the box: row blocks balanced by non-zero count, the x halo pulled by a kernel out of the owners' slices over
NVLink peer mappings while the interior rows are multiplied, the whole step one multi-device CUDA graph launch.  The
C++ twin is plugin/opt_b200.cpp built with -DB200_NGPU=N; the multi-process twin (torchrun, NCCL) is dist.py.
"""
import ctypes as C

import numpy as np

from ._lib import FORMATS, SYNTH, Options, check, lib
from .plugin import _ptr


class MgSpMat:
    def __init__(self, n_gpu, fmt="crs", **options):
        o = Options()
        for k, v in options.items():
            setattr(o, k, v)
        self.h = C.c_void_p()
        check(lib.b200spmv_mg_create(int(n_gpu), FORMATS[fmt], C.byref(o), C.byref(self.h)))
        self.nGPU = int(n_gpu)

    def convert_host(self, A):
        """A: plugin.SpMat (sorted COO on the host, square)."""
        check(lib.b200spmv_mg_convert_coo_host(self.h, A.nRow, A.nCol, A.nNnz, _ptr(A.row_idx), _ptr(A.col_idx), _ptr(A.val)))
        return self

    def convert_synth(self, kind, p0, p1=0, seed=1):
        check(lib.b200spmv_mg_convert_synth(self.h, SYNTH[kind], int(p0), int(p1), int(seed)))
        return self

    def scalar(self, name):
        v = C.c_longlong()
        check(lib.b200spmv_mg_get_scalar(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def bounds(self):
        b = np.empty(self.nGPU + 1, np.int32)
        check(lib.b200spmv_mg_get_bounds(self.h, _ptr(b)))
        return b

    def multiply_host(self, x, y):
        check(lib.b200spmv_mg_multiply_host(self.h, _ptr(x), _ptr(y)))

    def upload_x(self, x):
        check(lib.b200spmv_mg_upload_x(self.h, _ptr(x)))

    def multiply(self):
        check(lib.b200spmv_mg_multiply(self.h))

    def synchronize(self):
        check(lib.b200spmv_mg_synchronize(self.h))

    def download_y(self, y):
        check(lib.b200spmv_mg_download_y(self.h, _ptr(y)))

    def destroy(self):
        if self.h:
            lib.b200spmv_mg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
