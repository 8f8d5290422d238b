"""Probe (GPU): config 2 with one sliced-ELL matrix per column block (padding to the slice's longest row INSIDE the block), the
three multiplied one after the other -- what a padded column-blocked ELL engine would cost next to the tile-stream engine
(2.94 ms).  The running-sum hand-over between blocks (acc = y[r]) is not modelled (one more coalesced read of y per block)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
import singlespmv_b200 as sp
from singlespmv_b200._lib import lib, check


class Dev:
    def __init__(self, ptr, n, t):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": t, "data": (int(ptr), False), "version": 2}


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


n, k = (1 << 24, 32) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
d = sp.DeviceCoo("uniform", n, k, 1)
row = torch.as_tensor(Dev(d.c.row_d, d.nNnz, "<i4"), device="cuda")
col = torch.as_tensor(Dev(d.c.col_d, d.nNnz, "<i4"), device="cuda")
val = torch.as_tensor(Dev(d.c.val_d, d.nNnz, "<f8"), device="cuda")
x = torch.rand(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
for nb in (3, 4):
    B = (n + nb - 1) // nb
    mats, slots = [], 0
    for b in range(nb):
        m = (col >= b * B) & (col < (b + 1) * B)
        r, c, v = row[m].contiguous(), col[m].contiguous(), val[m].contiguous()
        del m
        A = sp.SpMatOpt("ell", col_blocks=-1)
        check(lib.b200spmv_convert_coo_device(A.h, n, n, r.numel(), C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(v.data_ptr()), None))
        slots += A.scalar("slots")
        mats.append(A)
        del r, c, v
    torch.cuda.empty_cache()

    def run():
        for A in mats:
            A.multiply(x.data_ptr(), y.data_ptr())
    ms = timed(run)
    print("uniform %d x %d, %d column blocks as sliced ELL: %.3f ms per multiply = %.1f GFLOP/s, slots / entries = %.3f, K per block %s"
          % (n, k, nb, ms, 2 * d.nNnz / ms / 1e6, slots / d.nNnz, [A.scalar("K") for A in mats]), flush=True)
    for A in mats:
        A.destroy()
