#!/usr/bin/env python
"""bench.py -- the reference's load -> convert -> repeat-multiply loop (src/main.cpp:58-102) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1..c5] [--format FMT] [--impl reference]

One "step" = one SpMV (y := A x) over the whole synthetic matrix of the workload.
Metric: SpMV GFLOP/s = 2 nnz / t (src/main.cpp:196), with effective HBM GB/s (compulsory bytes / t)
in `roofline`.  Workloads are BASELINE.json's configs:

    c1  2-D 5-point Laplacian 1024^2, CRS            c4  3-D 27-point 256^3, DIA
    c2  uniform random 2^24 rows x 32/row, CSS(3)    c5  3-D 7-point 512^3, row-partitioned CRS
    c3  R-MAT scale 23, 2^28 draws, CRS

N = 1 defaults to c2 (the config the metric is quoted on that fits one GPU); N > 1 defaults to c5, the
config BASELINE.json partitions over 1/2/4/8 GPUs (strong scaling, x halo exchanged every step).

--impl reference times the reference's own OpenMP CRS plugin (oracle/_ref/libref_crs.so, compiled
unmodified from /root/reference/src/opt_crs.cpp; else the C restatement in oracle/) on the host
cores, on a bounded row sample of the same workload.  That leg and `cpu_baseline` are the only
places this file touches oracle/.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    "c1": dict(kind="lap2d5", p0=1024, p1=0, seed=1, fmt="crs",
               name="CRS fp64, 2-D 5-point Laplacian 1024x1024 (1,048,576 rows, 5,238,784 nnz)"),
    "c2": dict(kind="uniform", p0=1 << 24, p1=32, seed=1, fmt="css", opts=dict(n_block=3), also=["ss", "ell", "jds"],
               name="SS family (column-blocked SS = CSS, N_BLOCK=3; SS / sliced-ELL / JDS in `formats`) fp64, "
                    "uniform random 16,777,216 rows x 32 nnz/row (536,870,912 nnz)"),
    "c3": dict(kind="rmat", p0=23, p1=1 << 28, seed=42, fmt="crs",
               name="adaptive CRS fp64, R-MAT scale 23, 2^28 edge draws (duplicates removed)"),
    "c4": dict(kind="box3d27", p0=256, p1=0, seed=1, fmt="dia",
               name="DIA fp64, 3-D 27-point stencil 256^3 (16,777,216 rows, 449,455,096 nnz)"),
    "c5": dict(kind="lap3d7", p0=512, p1=0, seed=1, fmt="crs",
               name="row-partitioned CRS fp64, 3-D 7-point Laplacian 512^3 (134,217,728 rows, 937,951,232 nnz)"),
}
MINI = {"c1": dict(p0=128), "c2": dict(p0=1 << 16), "c3": dict(p0=14, p1=1 << 18), "c4": dict(p0=32),
        "c5": dict(p0=64)}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu --set full
# captures (profiles/r1_ncu_kernels.md).  Keyed by (workload, format, n_block); anything else reports null.
NCU_TRAFFIC = {("c2", "css", 3): 2415011680, ("c2", "ell", 0): 37576549544, ("c2", "jds", 0): 33652983616,
               ("c3", "crs", 0): 3308867896, ("c3", "csr5", 0): 3383375304, ("c4", "dia", 0): 3870562096,
               ("c4", "ell", 0): 5853024560, ("c5", "crs", 0): 14787440000, ("c5", "dia", 0): 9634725000,
               ("c5", "csr5", 0): 13638726000, ("c5", "coo", 0): 20025184000}
DOMINANT = {"crs": "crs_rowblock_kernel (longest row <= 16) / tile_stream_kernel", "ss": "tile_stream_kernel", "css": "tile_stream_kernel (one launch per column block)",
            "ell": "ell_spmv_kernel", "jds": "jds_spmv_kernel", "dia": "dia_spmv_tma_kernel", "coo": "coo_tile_kernel",
            "csr5": "c5_compute_kernel"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU legs (oracle/)
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_crs(nRow, nCol, row, col, val, x, min_seconds, max_calls, warmup=1, exact_calls=None):
    """Times the reference's OpenMP CRS SpMV (src/opt_crs.cpp:44-70) on the host.  Returns
    (kind, seconds per call list).  Uses oracle/_ref when present, else the C restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib
    times = []
    # all host threads, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(cpu_threads()))
    except OSError:
        pass
    if oracle_lib.ref_available("crs"):
        kind = "reference"
        p = oracle_lib.RefPlugin("crs")
        lib = p.lib
        lib.ref_convert(C.c_int(nRow), C.c_int(nCol), C.c_int(len(row)), row.ctypes, col.ctypes, val.ctypes, x.ctypes)
        y = np.empty(nRow)
        call = lambda: lib.ref_spmv(y.ctypes)
    else:
        kind = "port"
        orc = oracle_lib.Oracle()
        m = orc.crs_convert(nRow, row, col, val)
        y = np.empty(nRow)
        lib = orc.lib
        call = lambda: lib.orc_crs_spmv(C.c_int(nRow), m["ptr"].ctypes, m["idx"].ctypes, m["val"].ctypes,
                                        x.ctypes, y.ctypes)
    for _ in range(warmup):
        call()
    t_begin = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        call()
        times.append(time.perf_counter() - t0)
        if exact_calls is not None:
            if len(times) >= exact_calls:
                break
        elif (time.perf_counter() - t_begin >= min_seconds and len(times) >= 3) or len(times) >= max_calls:
            break
    return kind, times


def sample_rows_for(wl, mini):
    """Row count of the bounded CPU sample: ~64 M non-zeros (a few 10 ms per call on a server CPU)."""
    kind, p0, p1 = wl["kind"], wl["p0"], wl["p1"]
    per_row = {"lap2d5": 5, "lap3d7": 7, "box3d27": 27, "uniform": p1, "rmat": 32}[kind]
    nRow = {"lap2d5": p0 * p0, "lap3d7": p0 ** 3, "box3d27": p0 ** 3, "uniform": p0, "rmat": 1 << p0}[kind]
    return min(nRow, max(1, (1 << 26) // per_row))


def host_sample(wl, rows):
    """The first `rows` rows of the workload's matrix, generated on the HOST by oracle/synth_oracle.c
    (bit-identical to the device generator, tests/test_gpu_parity.py::test_synth_matches_oracle)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib
    orc = oracle_lib.Oracle()
    kind, p0, p1, seed = wl["kind"], wl["p0"], wl["p1"], wl["seed"]
    if kind == "uniform":
        nRow, nCol, row, col, val = orc.uniform(seed, p0, p0, p1, 0, rows)
        nRow = rows
    elif kind == "rmat":
        nRow, nCol, row, col, val = orc.rmat(seed, p0, p1)
    else:
        nRow, nCol, row, col, val = orc.stencil_rows(kind, p0, 0, rows)
        nRow = rows
    x, _ = orc.reference_vectors(nCol, 0, 3)
    return nRow, nCol, row, col, val, x


# ------------------------------------------------------------------------------------------ reference arm
def run_reference_arm(args, wl, wl_key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(cpu_threads())        # before libgomp initialises
    rows = sample_rows_for(wl, args.mini)
    nRow, nCol, row, col, val, x = host_sample(wl, rows)
    kind, times = cpu_reference_crs(nRow, nCol, row, col, val, x, 0, 0, warmup=args.warmup, exact_calls=args.steps)
    t = sum(times) / len(times)
    gflops = 2.0 * len(row) / t / 1e9
    sample = "first %d rows (%d nnz) of the workload matrix, full-length x; reference CRS OpenMP SpMV" % (nRow, len(row))
    line = {"impl": "reference", "metric": "SpMV GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": wl_key + ": " + wl["name"], "format": "crs (reference src/opt_crs.cpp)",
                                            "sample": sample},
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cpu_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm, N = 1
def run_single(args, wl, wl_key):
    import numpy as np
    import torch
    import singlespmv_b200 as sp

    torch.cuda.set_device(0)
    fmt = args.format or wl["fmt"]
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    t0 = time.perf_counter()
    coo = sp.DeviceCoo(wl["kind"], wl["p0"], wl["p1"], wl["seed"])
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    options = dict(args.options)
    if fmt == "auto":                        # the engine's own pick from the matrix statistics (b200spmv_recommend_format)
        fmt, rec_opts = coo.recommend()
        for k, v in rec_opts.items():
            if not options.get(k):
                options[k] = v
    if not args.format:                      # the workload's own tunables apply to its default format only
        for k, v in wl.get("opts", {}).items():
            if not options.get(k):
                options[k] = v
    A = sp.SpMatOpt(fmt, **options).convert_device(coo)
    torch.cuda.synchronize()
    t_conv = time.perf_counter() - t0
    also = [] if (args.format or args.no_also) else list(wl.get("also", []))
    if not args.compare_cusparse and not also:
        coo.free()
    nRow, nCol, nnz = A.nRow, A.nCol, A.nNnz
    alg_bytes = A.scalar("alg_bytes")
    launches_per_step = A.scalar("launches")

    x_h, _ = sp.reference_vectors(nCol, 0, 3)                       # src/main.cpp:18,31
    x_pin = torch.from_numpy(x_h).pin_memory()
    y_pin = torch.empty(nRow, dtype=torch.float64).pin_memory()
    x_d = x_pin.cuda(non_blocking=True)
    y_d = torch.full((nRow,), float("nan"), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    need_flush = alg_bytes < 2 * L2_BYTES
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda") if need_flush else None

    def step():
        A.multiply(x_d.data_ptr(), y_d.data_ptr(), sptr)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    if need_flush:
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(args.steps):
            step()
        b.record(stream)
        torch.cuda.synchronize()
        total_ms = a.elapsed_time(b)
    warm_ms = None
    if need_flush:
        # for matrices that fit in L2 report the warm (L2-hot, back-to-back) figure as well
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            step()
        b.record(stream)
        torch.cuda.synchronize()
        warm_ms = a.elapsed_time(b) / args.steps
    # keep the sampler running over the e2e loop as well
    ms = total_ms / args.steps
    gflops = 2.0 * nnz / (ms * 1e-3) / 1e9

    # e2e: the reference-facing call with HOST vectors (SpMV(A_opt, x_opt, y)): H2D x, multiply, D2H y
    for _ in range(min(args.warmup, 3)):
        A.multiply_host(x_pin.numpy(), y_pin.numpy())
    xh, yh = x_pin.numpy(), y_pin.numpy()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        A.multiply_host(xh, yh)
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    assert torch.equal(torch.from_numpy(yh), y_d.cpu()), "host-semantics and device-resident results differ"

    peak, peak_src = peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    # the dominant kernel runs once per step, except CSS: once per column block (each streaming 1/nBlock of the matrix)
    dom_launches = A.scalar("nBlock") if fmt == "css" else 1
    traffic = None if args.mini else NCU_TRAFFIC.get((wl_key, fmt, options.get("n_block", 0) if fmt == "css" else 0))
    line = {"metric": "SpMV GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_key + ": " + wl["name"], "format": fmt, "options": {k: v for k, v in options.items() if v}, "nRow": nRow, "nCol": nCol, "nnz": nnz,
                       "x": "srand(3) rand()/RAND_MAX (src/main.cpp:18,31)",
                       "l2": "flushed between steps (512 MiB write)" if need_flush else "inputs larger than L2 (%.2f GB streamed per step)" % (alg_bytes / 1e9),
                       "convert_ms": t_conv * 1e3, "generate_ms": t_gen * 1e3,
                       **({"warm_l2_ms_per_step": warm_ms, "warm_l2_gflops": 2.0 * nnz / (warm_ms * 1e-3) / 1e9} if warm_ms else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": DOMINANT.get(fmt, fmt),
                         "alg_bytes_per_launch": alg_bytes // dom_launches, "dominant_launches_per_step": dom_launches,
                         "avg_launch_ms": ms / dom_launches,
                         "note": "achieved = alg_bytes_per_launch / avg_launch_ms (CUDA events over the timed region; the "
                                 "fix-up kernel's ~5 % share is inside); traffic = ncu dram read+write of one launch, "
                                 "profiles/r1_ncu_kernels.md"},
            "e2e": {"value": 2.0 * nnz / e2e_s / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": 8 * nCol, "d2h_bytes_per_step": 8 * nRow},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks}

    if also:
        # the other formats BASELINE.json names for this config: same matrix, same x, short device-resident runs
        line["formats"] = {fmt: {"gflops": gflops, "ms_per_step": ms, "frac": achieved / peak}}
        for f2 in also:
            B = sp.SpMatOpt(f2).convert_device(coo)
            for _ in range(3):
                B.multiply(x_d.data_ptr(), y_d.data_ptr(), sptr)
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n2 = max(3, args.steps // 5)
            ea.record(stream)
            for _ in range(n2):
                B.multiply(x_d.data_ptr(), y_d.data_ptr(), sptr)
            eb.record(stream)
            torch.cuda.synchronize()
            ms2 = ea.elapsed_time(eb) / n2
            line["formats"][f2] = {"gflops": 2.0 * nnz / (ms2 * 1e-3) / 1e9, "ms_per_step": ms2,
                                   "frac": B.scalar("alg_bytes") / (ms2 * 1e-3) / 1e9 / peak}
            B.destroy()
        if not args.compare_cusparse:
            coo.free()

    if args.compare_cusparse:
        # comparison point only (libb200cmp.so, singlespmv_b200/compare/): cusparseSpMV CSR on the same device arrays
        cmp = C.CDLL(os.path.join(ROOT, "singlespmv_b200", "libb200cmp.so"))
        cmp.b200cmp_last_error.restype = C.c_char_p
        res = {}
        y_c = torch.empty_like(y_d)
        for alg, name in ((1, "csr_alg1"), (2, "csr_alg2")):
            ms_c = C.c_float()
            st = cmp.b200cmp_cusparse_csr(C.c_int(nRow), C.c_int(nCol), C.c_longlong(nnz), C.c_void_p(coo.c.row_d),
                                          C.c_void_p(coo.c.col_d), C.c_void_p(coo.c.val_d), C.c_void_p(x_d.data_ptr()),
                                          C.c_void_p(y_c.data_ptr()), alg, args.warmup, args.steps, C.byref(ms_c), sptr)
            if st != 0:
                res[name] = {"error": cmp.b200cmp_last_error().decode()}
                continue
            torch.cuda.synchronize()
            rel = float(((y_c - y_d).abs().max() / y_d.abs().max()).item())
            res[name] = {"ms_per_step": ms_c.value, "gflops": 2.0 * nnz / (ms_c.value * 1e-3) / 1e9, "max_diff_rel_to_max_y": rel}
        line["cusparse"] = res
        coo.free()

    if not args.no_cpu:
        rows = sample_rows_for(wl, args.mini)
        if wl["kind"] == "rmat":
            c = sp.DeviceCoo(wl["kind"], wl["p0"], wl["p1"], wl["seed"])
        else:
            c = sp.DeviceCoo(wl["kind"], wl["p0"], wl["p1"], wl["seed"], 0, rows)
        _, _, row, col, val = c.to_host()
        c.free()
        s_rows = c.nRow if wl["kind"] == "rmat" else rows
        kind, times = cpu_reference_crs(s_rows, nCol, row, col, val, x_h, args.cpu_seconds, 2000)
        tbest = min(times)
        line["cpu_baseline"] = {"value": 2.0 * len(row) / tbest / 1e9, "unit": "GFLOP/s", "cores": cpu_threads(), "kind": kind,
                                "sample": "first %d rows (%d nnz) of the same matrix, full-length x; reference OpenMP CRS "
                                          "SpMV (src/opt_crs.cpp:44-70), min of %d calls over %.1f s (src/main.cpp:79-102 keeps the min)"
                                          % (s_rows, len(row), len(times), sum(times))}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--format", default=None)
    ap.add_argument("--mini", action="store_true", help="shrunken shapes (debugging only; not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the config's other formats")
    ap.add_argument("--no-graph", action="store_true", help="multi-GPU: launch each step eagerly instead of one CUDA graph")
    ap.add_argument("--compare-cusparse", action="store_true", help="also time cusparseSpMV CSR (comparison point)")
    ap.add_argument("--segment-width", type=int, default=0)
    ap.add_argument("--n-block", type=int, default=0)
    ap.add_argument("--sigma", type=int, default=0)
    ap.add_argument("--value-f32", action="store_true", help="CRS: fp32 storage of the matrix values, fp64 arithmetic")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.options = dict(segment_width=args.segment_width, n_block=args.n_block, csr5_sigma=args.sigma,
                        value_f32=1 if args.value_f32 else 0)
    wl_key = args.workload or ("c2" if args.gpus == 1 else "c5")
    wl = dict(WORKLOADS[wl_key])
    if args.mini:
        wl.update(MINI[wl_key])
        wl["name"] += " [MINI]"
    if args.impl == "reference":
        return run_reference_arm(args, wl, wl_key)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return run_single(args, wl, wl_key)
    from singlespmv_b200.dist import run_partitioned_bench
    return run_partitioned_bench(args, wl, wl_key)


if __name__ == "__main__":
    main()
