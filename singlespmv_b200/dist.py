"""Row-partitioned multi-GPU SpMV: one process per GPU, x halo over NCCL/NVLink.

The reference has no distributed code (single-node OpenMP; src/opt_crs.cpp:57-58 is its only "partition":
a static row schedule).  BASELINE.json config 5 asks for the CRS multiply row-partitioned by non-zero
balance with the x halo exchange overlapped with the local block.  Layering:

* device work (partition by nnz, halo discovery, column renumbering, pack kernel, multiply of a row range)
  is behind the C-ABI: b200spmv_partition_*, b200spmv_halo_*, b200spmv_multiply_rows;
* this file is host plumbing only: who asks whom for which x entries (``plan_requests``), and per multiply
  the exchange on a communication stream while the interior rows run on the compute stream.

Per multiply on every rank:
    comm stream   : the exchange.  Default ("peer"): ONE kernel over NVLink peer memory (csrc/xwin.cu) -- x_ext lives in
                    a window the other ranks map through CUDA IPC; the kernel signals, waits for the owners, pulls its
                    halo entries out of their slices and acknowledges.  Fallback ("nccl", B200SPMV_DIST_EXCHANGE=nccl or
                    when the windows cannot be mapped): pack kernel + grouped isend/irecv into x_ext's halo slots.
    compute stream: rows that touch no halo column            (b200spmv_multiply_rows, > 99 % of config 5)
    compute stream: after the halo has landed, the boundary rows at both ends of the block
"""
import ctypes as C

import numpy as np

from ._lib import SYNTH, Coo, check, lib
from .plugin import DeviceCoo, SpMatOpt, _ptr


# ------------------------------------------------------------------------------------------ host logic (no GPU needed)
def owner_counts(halo_cols, bounds):
    """How many of the (ascending, global) halo columns each block owns.  bounds: nParts+1 row/col splits."""
    halo_cols = np.asarray(halo_cols)
    edges = np.searchsorted(halo_cols, np.asarray(bounds), side="left")
    return np.diff(edges).astype(np.int64)


def plan_requests(halo_cols, bounds, rank, all_to_all_counts, all_to_all_lists):
    """Tell every owner which of its x entries this rank needs; learn what the peers need from us.

    all_to_all_counts(send_counts[nParts]) -> recv_counts[nParts]
    all_to_all_lists(send_list, send_counts, recv_counts) -> recv_list     (int32 global column ids)
    Returns (recv_counts = x entries arriving from each peer, send_counts = entries leaving to each peer,
             send_cols = global ids of owned columns to pack, in peer order)."""
    need = owner_counts(halo_cols, bounds)
    if need[rank] != 0:
        raise ValueError("halo of rank %d contains %d owned columns" % (rank, need[rank]))
    asked = np.asarray(all_to_all_counts(need), dtype=np.int64)
    send_cols = np.asarray(all_to_all_lists(np.ascontiguousarray(halo_cols, np.int32), need, asked), dtype=np.int32)
    lo, hi = bounds[rank], bounds[rank + 1]
    if len(send_cols) and (send_cols.min() < lo or send_cols.max() >= hi):
        raise ValueError("rank %d was asked for columns it does not own" % rank)
    return need, asked, send_cols


def host_bounds(row_idx, nRow, nParts):
    """Row split of a sorted host COO where the running non-zero count passes g nnz / nParts -- the rule of
    b200spmv_partition_rows (csrc/halo.cu) and b200spmv_mg_convert_coo_host, on the host."""
    nnz = len(row_idx)
    b = np.zeros(nParts + 1, np.int64)
    b[nParts] = nRow
    for g in range(1, nParts):
        e = nnz * g // nParts
        if e >= nnz:
            bnd = nRow
        else:
            r = int(row_idx[e])
            bnd = r if (e == 0 or int(row_idx[e - 1]) != r) else r + 1
        b[g] = max(bnd, b[g - 1])
    return b


class _DevArray:
    """A device pointer with the CUDA array interface, so that torch can view library-owned memory without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_array(ptr, n, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, n), device=device)


def plan_window(win, block, bounds, send_counts):
    """Pull plan of an x window: owners of the halo columns, and the ranks that read this one's slice."""
    readers = np.ascontiguousarray([p for p, c in enumerate(send_counts) if c], np.int32)
    b64 = np.ascontiguousarray(bounds, np.int64)
    hc = np.ascontiguousarray(block.halo_cols, np.int32)
    check(lib.b200spmv_xwin_plan(win, _ptr(hc) if len(hc) else None, len(hc), block.nLeft, block.nLocal, _ptr(b64),
                                 _ptr(readers) if len(readers) else None, len(readers)))


class Block:
    """One rank's share: local matrix (any format, default CRS), halo bookkeeping, device buffers.
    kind = a synthetic generator name, or a plugin.SpMat (host COO: the rank uploads only its own rows)."""

    def __init__(self, kind, p0, p1, seed, bounds, rank, fmt="crs", stream=None, crs_path=None):
        import os
        import torch
        self.rank, self.bounds = rank, [int(b) for b in bounds]
        rb, re = self.bounds[rank], self.bounds[rank + 1]
        if re <= rb:
            raise ValueError("rank %d owns no rows" % rank)
        coo = DeviceCoo(kind, p0, p1, seed, rb, re) if isinstance(kind, str) else DeviceCoo.from_host_rows(kind, rb, re)
        self.global_nnz_local = coo.nNnz
        self.halo = C.c_void_p()
        check(lib.b200spmv_halo_plan(C.byref(coo.c), rb, re, C.byref(self.halo), stream))
        info = (C.c_longlong * 8)()
        check(lib.b200spmv_halo_info(self.halo, info))
        self.nLocal, self.nLeft, self.nRight, self.interiorBegin, self.interiorEnd = (int(v) for v in info[:5])
        self.nRows = re - rb
        # crs_path 0 = the library's own choice (short rows: the TMA-fed row-chunk stream, a persistent kernel with one
        # CTA set per SM); 1 forces the tile-stream kernel, which round 1 used here because the row-block stream of that
        # round interleaved badly with the NCCL kernels (profiles/r1_experiments.md).  B200SPMV_DIST_CRS_PATH overrides.
        if crs_path is None:
            crs_path = int(os.environ.get("B200SPMV_DIST_CRS_PATH", "0"))
        self.A = SpMatOpt(fmt, crs_path=crs_path if fmt in ("crs", "ss") else 0).convert_device(coo, nRow=self.nRows, stream=stream)
        coo.free()
        # row-range bookkeeping now, so that no multiply_rows ever synchronises (they are captured in a CUDA graph)
        for rb_, re_ in ((self.interiorBegin, self.interiorEnd), (0, self.interiorBegin), (self.interiorEnd, self.nRows)):
            if re_ > rb_:
                self.A.prepare_rows(rb_, re_)
        n = lib.b200spmv_halo_cols(self.halo, None, 0)
        check(n)
        self.halo_cols = np.empty(n // 4, np.int32)
        if n:
            check(lib.b200spmv_halo_cols(self.halo, _ptr(self.halo_cols), n))
        self.x_ext = torch.zeros(self.nLeft + self.nLocal + self.nRight, dtype=torch.float64, device="cuda")
        self.y = torch.full((self.nRows,), float("nan"), dtype=torch.float64, device="cuda")
        self.sendbuf = None
        self.recv_counts = self.send_counts = None
        self.win = None

    @property
    def x_owned(self):
        return self.x_ext[self.nLeft:self.nLeft + self.nLocal]

    def use_window(self, win):
        """x_ext moves into an x window (csrc/xwin.cu) that the peers can map; call before set_requests."""
        import torch
        n = self.nLeft + self.nLocal + self.nRight
        self.win = win
        self.x_ext = device_array(lib.b200spmv_xwin_x_ext(win), n, self.x_ext.device) if n else self.x_ext
        self.x_ext.zero_()
        torch.cuda.synchronize()

    def set_requests(self, recv_counts, send_counts, send_cols):
        import torch
        self.recv_counts, self.send_counts = [int(v) for v in recv_counts], [int(v) for v in send_counts]
        send_cols = np.ascontiguousarray(send_cols, np.int32)
        check(lib.b200spmv_halo_set_send(self.halo, _ptr(send_cols), len(send_cols)))
        self.sendbuf = torch.empty(max(1, len(send_cols)), dtype=torch.float64, device="cuda")
        # views: where each peer's entries land in x_ext / leave from sendbuf
        self.recv_views, self.send_views = {}, {}
        off = 0
        for p, c in enumerate(self.recv_counts):
            if c:
                base = off if p < self.rank else self.nLocal + off     # left halo first, right halo after the owned slice
                self.recv_views[p] = self.x_ext[base:base + c]
            off += c
        off = 0
        for p, c in enumerate(self.send_counts):
            if c:
                self.send_views[p] = self.sendbuf[off:off + c]
            off += c

    def pack(self, stream_ptr=None):
        check(lib.b200spmv_halo_pack(self.halo, C.c_void_p(self.x_owned.data_ptr()), C.c_void_p(self.sendbuf.data_ptr()), stream_ptr))

    def multiply_interior(self, stream_ptr=None):
        if self.interiorEnd > self.interiorBegin:
            self.A.multiply_rows(self.interiorBegin, self.interiorEnd, self.x_ext.data_ptr(), self.y.data_ptr(), stream_ptr)

    def multiply_boundary(self, stream_ptr=None):
        if self.interiorBegin > 0:
            self.A.multiply_rows(0, self.interiorBegin, self.x_ext.data_ptr(), self.y.data_ptr(), stream_ptr)
        if self.interiorEnd < self.nRows:
            self.A.multiply_rows(self.interiorEnd, self.nRows, self.x_ext.data_ptr(), self.y.data_ptr(), stream_ptr)

    def launches_per_step(self):
        per = self.A.scalar("launches")
        parts = (self.interiorEnd > self.interiorBegin) + (self.interiorBegin > 0) + (self.interiorEnd < self.nRows)
        return per * parts + (1 if self.sendbuf is not None and sum(self.send_counts) else 0)

    def free(self):
        if self.win:
            self.x_ext = None
            lib.b200spmv_xwin_free(self.win)
            self.win = None
        if self.halo:
            lib.b200spmv_halo_free(self.halo)
            self.halo = C.c_void_p()
        self.A.destroy()


def synth_bounds(kind, p0, p1, nParts):
    b = np.empty(nParts + 1, np.int32)
    check(lib.b200spmv_partition_synth(SYNTH[kind], int(p0), int(p1), nParts, _ptr(b)))
    return b


# ------------------------------------------------------------------------------------------ all ranks in one process (tests)
def build_local_group(kind, p0, p1, seed, nParts, fmt="crs", windows=False):
    """All blocks on ONE GPU in one process: the same device code and the same request planning as the
    torchrun path (parity tests on a 1-GPU box).  windows=False: device-to-device copies stand in for NCCL send/recv;
    windows=True: every block's x_ext sits in an x window attached to the others (local_group_exchange_windows)."""
    bounds = synth_bounds(kind, p0, p1, nParts)
    blocks = [Block(kind, p0, p1, seed, bounds, r, fmt) for r in range(nParts)]
    need = [owner_counts(b.halo_cols, bounds) for b in blocks]
    if windows:
        wins = []
        for r, b in enumerate(blocks):
            w = C.c_void_p()
            check(lib.b200spmv_xwin_create(r, nParts, b.nLeft + b.nLocal + b.nRight, b.nLeft, C.byref(w)))
            wins.append(w)
        for r, b in enumerate(blocks):
            for p in range(nParts):
                if p != r:
                    check(lib.b200spmv_xwin_attach(wins[r], p, wins[p]))
            plan_window(wins[r], b, bounds, [need[p][r] for p in range(nParts)])
            b.use_window(wins[r])
    for r, b in enumerate(blocks):
        asked = np.array([need[p][r] for p in range(nParts)], np.int64)
        lists = []
        for p in range(nParts):
            off = int(need[p][:r].sum())
            lists.append(blocks[p].halo_cols[off:off + int(need[p][r])])
        send_cols = np.concatenate(lists) if lists else np.empty(0, np.int32)
        b.set_requests(need[r], asked, send_cols)
    return bounds, blocks


def local_group_multiply(blocks):
    for b in blocks:
        b.pack()
        b.multiply_interior()
    for b in blocks:
        for p, view in b.recv_views.items():
            view.copy_(blocks[p].send_views[b.rank])
    for b in blocks:
        b.multiply_boundary()


def local_group_multiply_windows(blocks, streams):
    """The torchrun step with all ranks in one process: every block's exchange kernel on its own stream (they wait for
    each other's flags, so they must be able to run at the same time), then the multiplies."""
    import torch
    cur = torch.cuda.current_stream()
    for b, st in zip(blocks, streams):
        st.wait_stream(cur)
        check(lib.b200spmv_xwin_exchange(b.win, C.c_void_p(st.cuda_stream)))
    for b in blocks:
        b.multiply_interior()
    for b, st in zip(blocks, streams):
        cur.wait_stream(st)
        b.multiply_boundary()


# ------------------------------------------------------------------------------------------ torch.distributed path
def _dist_plan(block, world, device):
    import torch
    import torch.distributed as dist

    def a2a_counts(send):
        s = torch.tensor(send, dtype=torch.int64, device=device)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s)
        return r.cpu().numpy()

    def a2a_lists(send_list, send_counts, recv_counts):
        s = torch.from_numpy(send_list).to(device)
        r = torch.empty(int(sum(recv_counts)), dtype=torch.int32, device=device)
        dist.all_to_all_single(r, s, [int(c) for c in recv_counts], [int(c) for c in send_counts])
        return r.cpu().numpy()

    return plan_requests(block.halo_cols, block.bounds, block.rank, a2a_counts, a2a_lists)


class DistSpmv:
    """The per-rank engine used by bench.py --gpus N (and usable as a library)."""

    def __init__(self, kind, p0, p1, seed, fmt="crs", exchange=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device("cuda", torch.cuda.current_device())
        # kind: generator name (every rank generates its own rows in HBM) or a plugin.SpMat every rank holds on the host
        self.bounds = synth_bounds(kind, p0, p1, self.world) if isinstance(kind, str) else host_bounds(kind.row_idx, kind.nRow, self.world)
        self.block = Block(kind, p0, p1, seed, self.bounds, self.rank, fmt)
        need, asked, send_cols = _dist_plan(self.block, self.world, self.device)
        import os
        self.exchange = exchange or os.environ.get("B200SPMV_DIST_EXCHANGE", "peer")
        self.exchange_error = None
        if self.exchange == "peer" and not self._map_windows(asked):
            self.exchange = "nccl"
        self.block.set_requests(need, asked, send_cols)
        self.compute = torch.cuda.current_stream()
        # highest priority: the tiny pack kernel (and NCCL's own kernels, TORCH_NCCL_HIGH_PRIORITY below) must not queue
        # behind the interior kernel's CTAs -- at 8 GPUs 0.063 of the 0.435 ms step was exposed exchange.  Not yet
        # re-measured (the round's GPU budget was spent); priorities cannot change results.
        self.comm = torch.cuda.Stream(priority=-1)
        self.graph, self.graph_error = None, None

    def _map_windows(self, send_counts):
        """x_ext into an x window, every peer's window mapped through CUDA IPC.  All ranks succeed or all fall back."""
        import torch
        import torch.distributed as dist
        b = self.block
        ok, win = 1, C.c_void_p()
        blob = np.zeros(80, np.uint8)
        try:
            check(lib.b200spmv_xwin_create(self.rank, self.world, b.nLeft + b.nLocal + b.nRight, b.nLeft, C.byref(win)))
            check(lib.b200spmv_xwin_export(win, _ptr(blob)))
        except Exception as e:                               # noqa: BLE001
            ok, self.exchange_error = 0, repr(e)
        mine = torch.from_numpy(blob).to(self.device)
        blobs = torch.empty(self.world * 80, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(blobs, mine)
        blobs = blobs.cpu().numpy().reshape(self.world, 80)
        if ok:
            try:
                for p in range(self.world):
                    if p != self.rank:
                        check(lib.b200spmv_xwin_import(win, p, _ptr(np.ascontiguousarray(blobs[p]))))
                plan_window(win, b, self.bounds, send_counts)
            except Exception as e:                           # noqa: BLE001
                ok, self.exchange_error = 0, repr(e)
        t = torch.tensor([ok], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if not int(t.item()):
            if win:
                lib.b200spmv_xwin_free(win)
            return False
        b.use_window(win)
        return True

    def _exchange(self):
        """The halo exchange on the comm stream (which the caller has already ordered after the writers of x_owned)."""
        import torch
        import torch.distributed as dist
        b = self.block
        if self.exchange == "peer":
            check(lib.b200spmv_xwin_exchange(b.win, C.c_void_p(self.comm.cuda_stream)))
            return
        with torch.cuda.stream(self.comm):
            b.pack(C.c_void_p(self.comm.cuda_stream))
            ops = [dist.P2POp(dist.irecv, v, p) for p, v in b.recv_views.items()]
            ops += [dist.P2POp(dist.isend, v, p) for p, v in b.send_views.items()]
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()                                     # comm stream waits for NCCL's stream

    def exchange_status(self):
        """(steps finished, timed_out) of the x-window exchange; (None, False) on the NCCL path.  Synchronises."""
        if self.exchange != "peer" or not self.block.win:
            return None, False
        steps, bad = C.c_longlong(), C.c_int()
        check(lib.b200spmv_xwin_status(self.block.win, C.byref(steps), C.byref(bad)))
        return int(steps.value), bool(bad.value)

    def _step_eager(self):
        """the exchange on the comm stream, interior rows on the current stream meanwhile, boundary rows once the halo
        has landed."""
        import torch
        b = self.block
        cur = torch.cuda.current_stream()
        cptr = C.c_void_p(cur.cuda_stream)
        self.comm.wait_stream(cur)                           # previous readers of x_ext / writers of x_owned are ordered before
        self._exchange()
        b.multiply_interior(cptr)                            # overlaps the exchange
        cur.wait_stream(self.comm)
        b.multiply_boundary(cptr)

    # ---- host-semantics step: x_owned comes from a pinned host slice, y goes back to one, inside the step
    def plan_host(self, max_chunks=8):
        """Row chunks of the interior and the piece of x each needs (b200spmv_rows_col_extent): a banded block starts
        multiplying while the rest of its x slice is still crossing PCIe, and y flows back chunk by chunk."""
        import torch
        b = self.block
        ib, ie = b.interiorBegin, b.interiorEnd
        n = max(1, min(max_chunks, (ie - ib) * 8 // (4 << 20)))
        cuts = [ib + (((ie - ib) * c // n) & ~31) for c in range(n)] + [ie]
        self.h_chunks = [(cuts[c], cuts[c + 1]) for c in range(n) if cuts[c + 1] > cuts[c]]
        self.h_pieces = [((b.nLocal * k // n) & ~31) for k in range(n)] + [b.nLocal]
        self.h_need = []
        for rb, re in self.h_chunks:
            lo, hi = b.A.col_extent(rb, re)
            own = hi - b.nLeft                                   # local numbering: [left halo | owned | right halo]
            k = 0
            while k + 1 < n and self.h_pieces[k + 1] <= own:
                k += 1
            self.h_need.append(k if hi >= lo else -1)
        self.h_in, self.h_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.h_ev_in = [torch.cuda.Event() for _ in range(n)]
        self.h_ev_out = [torch.cuda.Event() for _ in range(len(self.h_chunks) + 1)]

    def step_host(self, x_pin, y_pin):
        """One multiply with host vectors: H2D of the owned x slice (pieces), exchange + multiply, D2H of the owned
        y slice (chunks) -- all three overlapped; returns without synchronising (streams h_out / current hold the tail)."""
        import torch
        b = self.block
        cur = torch.cuda.current_stream()
        cptr = C.c_void_p(cur.cuda_stream)
        self.h_in.wait_stream(cur)                               # the previous step's readers of x_ext
        cur.wait_stream(self.h_out)                              # the previous step's D2H of y
        n = len(self.h_ev_in)
        with torch.cuda.stream(self.h_in):
            for k in range(n):
                p0, p1 = self.h_pieces[k], self.h_pieces[k + 1]
                if p1 > p0:
                    b.x_owned[p0:p1].copy_(x_pin[p0:p1], non_blocking=True)
                self.h_ev_in[k].record(self.h_in)
        self.comm.wait_stream(cur)
        self.comm.wait_event(self.h_ev_in[-1])                   # the peers read both ends of the slice
        self._exchange()
        for i, (rb, re) in enumerate(self.h_chunks):
            if self.h_need[i] >= 0:
                cur.wait_event(self.h_ev_in[self.h_need[i]])
            b.A.multiply_rows(rb, re, b.x_ext.data_ptr(), b.y.data_ptr(), cptr)
            self.h_ev_out[i].record(cur)
            self.h_out.wait_event(self.h_ev_out[i])
            with torch.cuda.stream(self.h_out):
                y_pin[rb:re].copy_(b.y[rb:re], non_blocking=True)
        cur.wait_event(self.h_ev_in[-1])
        cur.wait_stream(self.comm)
        b.multiply_boundary(cptr)
        self.h_ev_out[-1].record(cur)
        self.h_out.wait_event(self.h_ev_out[-1])
        with torch.cuda.stream(self.h_out):
            if b.interiorBegin > 0:
                y_pin[:b.interiorBegin].copy_(b.y[:b.interiorBegin], non_blocking=True)
            if b.interiorEnd < b.nRows:
                y_pin[b.interiorEnd:].copy_(b.y[b.interiorEnd:], non_blocking=True)

    def enable_graph(self, warmup=3):
        """Capture one step (both streams, the NCCL send/recv included) in a CUDA graph: one launch per step
        instead of ~8 kernel launches + a grouped NCCL call from Python.  Falls back to eager on any failure."""
        import torch
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._step_eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:                               # noqa: BLE001
            self.graph, self.graph_error = None, repr(e)
            try:
                torch.cuda.synchronize()
            except Exception:                                # noqa: BLE001
                pass
        return self.graph is not None

    def choose_launch(self, steps=10):
        """Graph replay or eager launches, whichever is faster for this block size (max over ranks, decided together):
        a step of ~1 ms hides the eager launches completely and the graph's fork/join nodes only add latency (2 GPUs:
        1.029 ms eager, 1.052 ms replayed); at 0.3 ms per step (8 GPUs) the single launch wins."""
        import torch
        import torch.distributed as dist
        if self.graph is None:
            return "eager"
        g = self.graph

        def timed(use_graph):
            self.graph = g if use_graph else None
            for _ in range(3):
                self.step()
            torch.cuda.synchronize()
            dist.barrier()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                self.step()
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(e)], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        t_graph, t_eager = timed(True), timed(False)
        self.graph = g if t_graph <= t_eager else None
        self.kept_graph = g                                  # stays alive either way (released in release_graph)
        return "graph" if self.graph is not None else "eager"

    def release_graph(self):
        import gc
        self.graph = None
        self.kept_graph = None
        gc.collect()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_eager()


def bind_to_gpu_numa_node(local_rank):
    """Run this process (and allocate its page-locked vectors, first touch) on the CPUs NVML reports as closest to the
    GPU: 8 ranks staging 2 x 1 GB per step through one socket's memory was the e2e limit (15 ms at 2 and at 8 GPUs).
    Returns the number of CPUs bound to, or None when NVML / the affinity call is not available."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:                                        # noqa: BLE001
        return None


def run_partitioned_bench(args, wl, wl_key):
    """bench.py --gpus N (N > 1): strong scaling of one matrix over N ranks, launched by torchrun."""
    import json
    import os
    import sys
    import time
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, CpuReferenceCrs, parity_of, peaks, sample_rows_for

    # NCCL prints its version banner on stdout; the contract is ONE JSON line there -> park fd 1 on stderr meanwhile
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")     # NCCL's internal streams above the compute stream
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    fmt = args.format or "crs"
    eng = DistSpmv(wl["kind"], wl["p0"], wl["p1"], wl["seed"], fmt)
    b = eng.block
    nRow = int(eng.bounds[-1])
    # x: every rank fills its owned slice from the reference's rand() stream (src/main.cpp:18,31)
    from .plugin import reference_vectors
    x_h, _ = reference_vectors(nRow, 0, 3)
    lo, hi = int(eng.bounds[rank]), int(eng.bounds[rank + 1])
    numa = bind_to_gpu_numa_node(local_rank)                   # page-locked vectors on the memory next to this GPU's PCIe root
    x_pin = torch.from_numpy(x_h[lo:hi].copy()).pin_memory()
    y_pin = torch.empty(b.nRows, dtype=torch.float64).pin_memory()
    del x_h
    b.x_owned.copy_(x_pin, non_blocking=True)
    torch.cuda.synchronize()

    nnz_t = torch.tensor([b.global_nnz_local, b.A.scalar("alg_bytes"), b.nLeft + b.nRight], dtype=torch.int64, device="cuda")
    dist.all_reduce(nnz_t)
    nnz, alg_bytes_sum, halo_total = (int(v) for v in nnz_t.tolist())
    # compulsory bytes of the GLOBAL multiply (SURVEY.md 8d CRS formula), not the sum of the local ones
    alg_bytes = 12 * nnz + 4 * (nRow + 1) + 8 * nRow + 8 * nRow

    graphed = False
    if not getattr(args, "no_graph", False):
        ok = torch.tensor([1 if eng.enable_graph() else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)            # all ranks graphed, or none
        graphed = bool(ok.item())
        if not graphed:
            eng.graph = None
        else:
            graphed = eng.choose_launch() == "graph"
    for _ in range(args.warmup):
        eng.step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    a.record(eng.compute)
    for _ in range(args.steps):
        eng.step()
    e.record(eng.compute)
    torch.cuda.synchronize()
    ms_local = a.elapsed_time(e)
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    dist.barrier()

    # exposed communication: eager full step vs the same launches without the exchange
    def compute_only():
        cptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        b.multiply_interior(cptr)
        b.multiply_boundary(cptr)

    def timed(fn):
        torch.cuda.synchronize()
        dist.barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(eng.compute)
        for _ in range(args.steps):
            fn()
        eb.record(eng.compute)
        torch.cuda.synchronize()
        tt = torch.tensor([ea.elapsed_time(eb) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())
    eager_full_ms = timed(eng._step_eager)
    compute_only_ms = timed(compute_only)

    # sampled full-size parity: the first rows of the matrix (all inside rank 0's block) against the reference's CRS
    parity = None
    if rank == 0 and not getattr(args, "no_cpu", False):
        import numpy as np
        rows = min(sample_rows_for(wl), b.nRows)
        c = DeviceCoo(wl["kind"], wl["p0"], wl["p1"], wl["seed"], 0, rows)
        _, _, row, col, val = c.to_host()
        c.free()
        x_full, _ = reference_vectors(nRow, 0, 3)
        ref = CpuReferenceCrs(rows, nRow, row, col, val, x_full)
        ref.call()
        mag = np.bincount(row, weights=np.abs(val * x_full[col]), minlength=rows)[:rows]
        parity = dict(parity_of(b.y[:rows].cpu().numpy(), ref.y, mag),
                      against="reference CRS (src/opt_crs.cpp:44-70) on the first rows of the same matrix (rank 0's block)")
        del ref, row, col, val, x_full, mag
    dist.barrier()

    # e2e: per step H2D of the owned x slice, exchange + multiply, D2H of the owned y slice -- the three overlapped
    # (x in pieces, y in chunks; DistSpmv.step_host)
    eng.plan_host()

    def e2e_step():
        eng.step_host(x_pin, y_pin)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    assert torch.equal(y_pin, b.y.cpu()), "host-semantics and device-resident results differ"
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    e2e_local = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = torch.tensor([b.launches_per_step()], dtype=torch.int64, device="cuda")
    dist.all_reduce(launches)
    bad = torch.tensor([1 if eng.exchange_status()[1] else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(bad)

    if rank == 0:
        peak, peak_src = peaks()
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        line = {"metric": "SpMV GFLOP/s", "value": 2.0 * nnz / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl_key + ": " + wl["name"], "format": fmt, "nRow": nRow, "nCol": nRow, "nnz": nnz,
                           "parallelism": "row blocks by nnz balance x%d, x halo %d doubles/step %s, overlapped with interior rows"
                                          % (world, halo_total, "pulled out of the owners' x windows over NVLink peer memory "
                                             "(CUDA IPC) by one exchange kernel per rank" if eng.exchange == "peer" else
                                             "over NCCL send/recv"),
                           "exchange": eng.exchange, "exchange_fallback_reason": eng.exchange_error,
                           "exchange_flag_timeouts": int(bad.item()),
                           "local_kernel": ("crs_tma_kernel (row-chunk stream)" if fmt == "crs" and b.A.scalar("short_row_path") else
                                            "tile_stream_kernel" if fmt in ("crs", "ss", "css") else fmt),
                           "launch": ("one CUDA graph per step (both streams captured)" if graphed else "eager launches") + "; the faster of the two, timed at plan time",
                           "eager_ms_per_step": eager_full_ms, "compute_only_ms_per_step": compute_only_ms,
                           "exposed_exchange_ms": max(0.0, eager_full_ms - compute_only_ms),
                           "l2": "inputs larger than L2 (%.2f GB per GPU per step)" % (alg_bytes / world / 1e9)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                             "frac": achieved / (peak * world), "traffic": None,
                             "peak_source": peak_src + " x %d GPUs" % world, "alg_bytes_per_launch": alg_bytes,
                             "kernel": "CRS multiply (crs_tma_kernel for short rows, else tile_stream_kernel): whole step incl. exposed halo exchange"},
                "e2e": {"value": 2.0 * nnz / e2e_s / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_s * 1e3,
                        "h2d_bytes_per_step": 8 * nRow, "d2h_bytes_per_step": 8 * nRow, "host_cpus_bound": numa},
                "gpu_launches": int(launches.item()) * args.steps, "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    eng.release_graph()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    # a captured NCCL graph can keep ncclCommDestroy waiting at interpreter teardown (seen on 2 x B200, torch 2.11 /
    # NCCL 2.28.9): everything is flushed and every rank has passed the barrier, so leave without the destructors
    os._exit(0)
