"""Multi-GPU check of the single-process path (b200spmv_mg_*), run directly on a box with >= 2 GPUs:

    python tests/mg_check.py [nGPU]

Runs the mg parity tests of tests/test_gpu_parity.py (they use every GPU count the box offers), then config 5 through the
C-ABI at full size: y of the partitioned multiply against the single-GPU CRS result (bit-identical) and the device-
resident step time (one multi-device CUDA graph launch per step)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import singlespmv_b200 as sp                      # noqa: E402
from singlespmv_b200.mg import MgSpMat            # noqa: E402


def main():
    import torch
    n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else sp.device_count()
    p0 = int(os.environ.get("MG_CHECK_P0", "512"))
    M = MgSpMat(n_gpu, "crs").convert_synth("lap3d7", p0)
    n, nnz = M.scalar("nRow"), M.scalar("nNnz")
    if os.environ.get("MG_CHECK_NO_SINGLE"):
        # a matrix one GPU cannot hold (more than 2^31-1 non-zeros in all; every block below that): A.1 = row sums, closed form
        x = np.ones(n)
        y = np.full(n, np.nan)
        M.multiply_host(x, y)
        r = np.arange(n, dtype=np.int64)

        def span(i):
            return 1 + (i > 0).astype(np.int64) + (i < p0 - 1).astype(np.int64)
        cnt = span(r // (p0 * p0)) + span((r // p0) % p0) + span(r % p0) - 2
        same = bool(np.array_equal(y, (7 - cnt).astype(np.float64)))
        assert nnz == int(cnt.sum())
        del r, cnt
    else:
        x, _ = sp.reference_vectors(n, 0, 3)
        y = np.full(n, np.nan)
        M.multiply_host(x, y)
        torch.cuda.set_device(0)
        coo = sp.DeviceCoo("lap3d7", p0)
        A = sp.SpMatOpt("crs").convert_device(coo)
        coo.free()
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty(n, dtype=torch.float64, device="cuda")
        A.multiply(xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        same = bool(np.array_equal(y, yd.cpu().numpy()))
        A.destroy()
        del xd, yd
    M.upload_x(x)
    for _ in range(5):
        M.multiply()
    M.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(50):
            M.multiply()
        M.synchronize()
        best = min(best, (time.perf_counter() - t0) / 50)
    # host semantics as the C++ plugin issues them: the caller's vectors page-locked once (opt_b200.cpp does it in OptimizeProblem)
    import ctypes as C
    from singlespmv_b200._lib import lib
    for v in (x, y):
        lib.b200spmv_host_register(v.ctypes.data_as(C.c_void_p), C.c_ulonglong(v.nbytes))
    M.multiply_host(x, y)
    t0 = time.perf_counter()
    for _ in range(5):
        M.multiply_host(x, y)
    host = (time.perf_counter() - t0) / 5
    for v in (x, y):
        lib.b200spmv_host_unregister(v.ctypes.data_as(C.c_void_p))
    print("mg_check lap3d7 p0=%d nGPU=%d rows=%d nnz=%d bit-identical=%s graphed=%d halo=%d step %.4f ms = %.1f GFLOP/s (frac of %d x 6543 GB/s: %.3f) host-semantics %.2f ms"
          % (p0, n_gpu, n, nnz, same, M.scalar("graphed"), M.scalar("halo_total"), best * 1e3, 2.0 * nnz / best / 1e9, n_gpu,
             M.scalar("alg_bytes") / best / 1e9 / (6543.4 * n_gpu), host * 1e3), flush=True)
    M.destroy()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
