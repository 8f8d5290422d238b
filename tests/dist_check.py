"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank multiplies its row block with both halo exchanges of singlespmv_b200.dist.DistSpmv (the peer-memory x
windows of csrc/xwin.cu and the NCCL send/recv fallback); rank 0
also multiplies the whole matrix on its own GPU with the single-GPU CRS path.  The blocks are renumbered
monotonically, so the concatenated y must be bit-identical."""
import os
import sys

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import singlespmv_b200 as sp                      # noqa: E402
from singlespmv_b200._lib import lib              # noqa: E402
from singlespmv_b200.dist import DistSpmv         # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    cases = (("lap3d7", 96, 0), ("box3d27", 48, 0), ("lap2d5", 700, 0), ("uniform", 1 << 16, 16), ("rmat-host", 14, 1 << 18))
    for kind, p0, p1, exchange in [c + (e,) for c in cases for e in ("peer", "nccl")]:
        host = None
        if kind == "rmat-host":        # a matrix that arrives as a host COO (e.g. from a .mtx file): every rank uploads only its rows
            d = sp.DeviceCoo("rmat", p0, p1, 42)
            host = sp.SpMat(*d.to_host())
            d.free()
        eng = DistSpmv(host if host is not None else kind, p0, p1, 1, exchange=exchange)
        n = int(eng.bounds[-1])
        x_h, _ = sp.reference_vectors(n, 0, 3)
        lo, hi = int(eng.bounds[rank]), int(eng.bounds[rank + 1])
        eng.block.x_owned.copy_(torch.from_numpy(x_h[lo:hi]))
        for _ in range(3):                       # repeated steps must be idempotent
            eng.step()
        graphed = eng.enable_graph() if os.environ.get("DIST_CHECK_GRAPH", "1") == "1" else False
        eng.block.y.fill_(float("nan"))
        for _ in range(2):
            eng.step()
        torch.cuda.synchronize()
        # host-semantics step (x slice from pinned host memory in pieces, y back in chunks): same bits as the device step
        y_dev = eng.block.y.clone()
        x_pin = torch.from_numpy(x_h[lo:hi].copy()).pin_memory()
        y_pin = torch.full((hi - lo,), float("nan"), dtype=torch.float64).pin_memory()
        eng.plan_host(max_chunks=5)
        eng.block.x_owned.zero_()
        for _ in range(2):
            eng.step_host(x_pin, y_pin)
        torch.cuda.synchronize()
        host_ok = bool(torch.equal(y_pin, y_dev.cpu()))
        # a NEW x every step must reach the peers: x doubled -> y doubled exactly, eagerly and through the graph
        for _ in range(2):
            eng.block.x_owned.mul_(2.0)
            eng.step()
        torch.cuda.synchronize()
        host_ok = host_ok and bool(torch.equal(eng.block.y, 4.0 * y_dev))
        eng.block.x_owned.mul_(0.25)
        eng.step()
        torch.cuda.synchronize()
        host_ok = host_ok and bool(torch.equal(eng.block.y, y_dev))
        if eng.block.win:
            steps, bad = C.c_longlong(), C.c_int()
            lib.b200spmv_xwin_status(eng.block.win, C.byref(steps), C.byref(bad))
            host_ok = host_ok and bad.value == 0
        flag_h = torch.tensor([1 if host_ok else 0], device="cuda")
        dist.all_reduce(flag_h, op=dist.ReduceOp.MIN)
        host_ok = bool(flag_h.item())
        sizes = [int(eng.bounds[r + 1] - eng.bounds[r]) for r in range(world)]
        parts = [torch.empty(s, dtype=torch.float64, device="cuda") for s in sizes]
        dist.all_gather(parts, eng.block.y)
        if rank == 0:
            y = torch.cat(parts)
            coo = sp.DeviceCoo(kind, p0, p1, 1) if host is None else sp.DeviceCoo.from_host_rows(host, 0, host.nRow)
            A = sp.SpMatOpt("crs").convert_device(coo)
            xd = torch.from_numpy(x_h).cuda()
            y1 = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
            A.multiply(xd.data_ptr(), y1.data_ptr())
            torch.cuda.synchronize()
            same = bool(torch.equal(y, y1)) if host is None else bool(torch.allclose(y, y1, rtol=1e-12, atol=1e-300))   # warp-reduced long rows
            halo = eng.block.nLeft + eng.block.nRight
            print("dist_check %s p0=%d world=%d exchange=%s%s rows=%d halo(rank0)=%d interior(rank0)=[%d,%d) bit-identical=%s host-step+fresh-x=%s graph=%s %s"
                  % (kind, p0, world, eng.exchange, "" if eng.exchange == exchange else " (FELL BACK: %s)" % eng.exchange_error, n, halo,
                     eng.block.interiorBegin, eng.block.interiorEnd, same, host_ok, graphed, eng.graph_error or ""), flush=True)
            ok = ok and same and host_ok
        dist.barrier()
        eng.release_graph()
        torch.cuda.synchronize()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    code = 0 if int(flag.item()) == 1 else 1
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)        # see singlespmv_b200/dist.py: teardown with captured NCCL graphs can hang


if __name__ == "__main__":
    main()
