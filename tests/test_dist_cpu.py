"""The N > 1 path's host logic on CPU: world_size-2 (and 3) gloo process groups run the same request
planning (singlespmv_b200.dist.plan_requests / _dist_plan) the NCCL path uses, exchange x entries with
all_to_all, and the assembled [left halo | owned | right halo] vector must reproduce the global product."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nnz_bounds(ptr, nParts):
    nnz = int(ptr[-1])
    b = [int(np.searchsorted(ptr, nnz * g // nParts, side="left")) for g in range(nParts + 1)]
    b[0], b[-1] = 0, len(ptr) - 1
    return b


def _worker(rank, world, port, kind, n):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle
    from singlespmv_b200 import dist as spd

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        nRow, nCol, row, col, val = orc.stencil(kind, n)
        x, _ = orc.reference_vectors(nCol, nRow)
        y_ref = orc.crs_result(nRow, row, col, val, x)
        ptr = orc.crs_convert(nRow, row, col, val)["ptr"]
        bounds = _nnz_bounds(ptr, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        _, _, r, c, v = orc.stencil_rows(kind, n, lo, hi)
        halo = np.unique(c[(c < lo) | (c >= hi)]).astype(np.int32)

        blk = type("Blk", (), {"halo_cols": halo, "rank": rank, "bounds": bounds})()
        need, asked, send_cols = spd._dist_plan(blk, world, torch.device("cpu"))
        assert need.sum() == len(halo) and need[rank] == 0 and asked[rank] == 0
        assert np.all((send_cols >= lo) & (send_cols < hi))
        # exchange the values and assemble x_ext
        sendbuf = torch.from_numpy(x[send_cols])
        recvbuf = torch.empty(int(need.sum()), dtype=torch.float64)
        dist.all_to_all_single(recvbuf, sendbuf, [int(t) for t in need], [int(t) for t in asked])
        nLeft = int((halo < lo).sum())
        x_ext = np.concatenate([recvbuf.numpy()[:nLeft], x[lo:hi], recvbuf.numpy()[nLeft:]])
        assert np.array_equal(recvbuf.numpy(), x[halo])
        # monotone local numbering (what b200spmv_halo_plan does on the device)
        k = np.searchsorted(halo, c)
        lc = np.where(c < lo, k, np.where(c >= hi, (hi - lo) + k, nLeft + (c - lo))).astype(np.int32)
        assert np.all(np.diff(lc.astype(np.int64) + (r.astype(np.int64) - lo) * len(x_ext)) > 0)   # still sorted
        y = orc.crs_result(hi - lo, r - lo, lc, v, x_ext)
        assert np.array_equal(y, y_ref[lo:hi])           # same summation order -> bit-identical
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,n", [(2, "lap3d7", 10), (3, "box3d27", 7), (2, "lap2d5", 25)])
def test_partitioned_plan_gloo(world, kind, n):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), kind, n), nprocs=world, join=True)


def test_owner_counts():
    from singlespmv_b200.dist import owner_counts
    assert owner_counts([0, 1, 5, 9, 10], [0, 4, 8, 12]).tolist() == [2, 1, 2]
    assert owner_counts([], [0, 4, 8]).tolist() == [0, 0]


def test_host_bounds_rule():
    """host_bounds = the split rule of b200spmv_partition_rows / b200spmv_mg_convert_coo_host on a host COO: block g starts at
    the row holding entry g nnz / G (the next row when that entry is not the row's first); monotone, covers all rows."""
    from singlespmv_b200.dist import host_bounds
    rng = np.random.default_rng(1)
    for nRow, parts in ((10, 3), (1000, 8), (7, 7), (50, 2)):
        lens = rng.integers(0, 9, nRow)
        lens[rng.integers(0, nRow)] = 300                       # one heavy row
        row = np.repeat(np.arange(nRow), lens).astype(np.int32)
        b = host_bounds(row, nRow, parts)
        assert b[0] == 0 and b[-1] == nRow and np.all(np.diff(b) >= 0)
        ptr = np.concatenate([[0], np.cumsum(lens)])
        for g in range(1, parts):
            e = len(row) * g // parts
            assert ptr[b[g]] >= e or b[g] == nRow             # the block starts at or after the balance point ...
            if b[g] > 0 and b[g] > b[g - 1]:
                assert ptr[b[g] - 1] <= e                     # ... and no earlier row boundary would have done
    assert list(host_bounds(np.zeros(0, np.int32), 4, 2)) == [0, 4, 4]
