"""Probe (GPU): one L2-resident column block of config 2 -- uniform random, x = 45 MB -- through the COO entry stream (load-fed)
and through the CRS tile-stream the column-block engine uses today.  3 x the time here ~ a column-blocked multiply of config 2."""
import sys
import torch
sys.path.insert(0, ".")
import singlespmv_b200 as sp

def timed(fn, steps=20, warm=5):
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

for n, k in ((5600000, 32), (5600000, 11), (16777216, 11)):
    d = sp.DeviceCoo("uniform", n, k, 1)
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for fmt, opt in (("crs", {}), ("coo", {"coo_path": 2}), ("coo", {"coo_path": 3}), ("ell", {"col_blocks": -1})):
        A = sp.SpMatOpt(fmt, **opt).convert_device(d)
        ms = timed(lambda: A.multiply(x.data_ptr(), y.data_ptr()))
        print("uniform n=%d K=%d x=%.0f MB nnz=%d %s %s: %.4f ms = %.1f G entries/s, %.1f GFLOP/s" % (n, k, n * 8 / 2**20, d.nNnz, fmt, opt, ms, d.nNnz / ms / 1e6, 2 * d.nNnz / ms / 1e6), flush=True)
        A.destroy()
    d.free()
