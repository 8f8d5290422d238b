"""ctypes bindings for the TEST-ONLY checkers under oracle/.

* ``Oracle``   -- oracle/liboracle.so, our C restatement of the reference algorithms
                 (oracle/spmv_oracle.c, synth_oracle.c, csr5_oracle.c).
* ``RefPlugin`` -- oracle/_ref/libref_<variant>.so, the reference's own plugin compiled unmodified
                 (oracle/ref_shim.cpp + /root/reference/src/*.cpp via oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_oracle(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith("_oracle.c")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", ORACLE_DIR, "ref"])
    return so


class Oracle:
    """The C restatement.  Inputs: COO sorted by (row, col), no duplicates."""

    def __init__(self):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        self.lib = C.CDLL(so)
        L = self.lib
        L.synth_stencil_rows.restype = C.c_longlong
        L.synth_stencil_nnz.restype = C.c_longlong
        L.synth_stencil.restype = C.c_longlong
        L.synth_rmat.restype = C.c_longlong
        L.synth_stencil_range.restype = C.c_longlong

    # ---- CRS (reference src/opt_crs.cpp)
    def crs_convert(self, nRow, row, col, val):
        row = _i32(row)
        ptr = np.empty(nRow + 1, np.int32)
        self.lib.orc_crs_row_ptr(C.c_int(nRow), C.c_int(len(row)), row.ctypes, ptr.ctypes)
        return {"ptr": ptr, "idx": _i32(col).copy(), "val": _f64(val).copy()}

    def crs_spmv(self, m, x, nRow=None):
        nRow = len(m["ptr"]) - 1 if nRow is None else nRow
        y = np.full(nRow, np.nan)
        x = _f64(x)
        self.lib.orc_crs_spmv(C.c_int(nRow), m["ptr"].ctypes, m["idx"].ctypes, m["val"].ctypes,
                              x.ctypes, y.ctypes)
        return y

    def crs_result(self, nRow, row, col, val, x):
        """The results oracle: y of the reference CRS plugin."""
        return self.crs_spmv(self.crs_convert(nRow, row, col, val), x)

    # ---- COO (reference src/opt_coo.cpp)
    def coo_spmv(self, nRow, row, col, val, x):
        row, col, val, x = _i32(row), _i32(col), _f64(val), _f64(x)
        y = np.full(nRow, np.nan)
        self.lib.orc_coo_spmv(C.c_int(nRow), C.c_int(len(row)), row.ctypes, col.ctypes, val.ctypes,
                              x.ctypes, y.ctypes)
        return y

    # ---- ELL (reference src/opt_ell.cpp)
    def ell_convert(self, nRow, row, col, val):
        row, col, val = _i32(row), _i32(col), _f64(val)
        K = self.lib.orc_ell_width(C.c_int(nRow), C.c_int(len(row)), row.ctypes)
        ecol = np.empty((nRow, K), np.int32)
        evalv = np.empty((nRow, K), np.float64)
        self.lib.orc_ell_convert(C.c_int(nRow), C.c_int(len(row)), C.c_int(K), row.ctypes, col.ctypes,
                                 val.ctypes, ecol.ctypes, evalv.ctypes)
        return {"K": K, "col_idx": ecol, "val": evalv}

    def ell_spmv(self, m, x):
        nRow = m["col_idx"].shape[0]
        y = np.full(nRow, np.nan)
        x = _f64(x)
        self.lib.orc_ell_spmv(C.c_int(nRow), C.c_int(m["K"]), m["col_idx"].ctypes, m["val"].ctypes,
                              x.ctypes, y.ctypes)
        return y

    # ---- JDS (reference src/opt_jds.cpp)
    def jds_convert(self, nRow, row, col, val, perm_in=None):
        row, col, val = _i32(row), _i32(col), _f64(val)
        nnz = len(row)
        K = self.lib.orc_ell_width(C.c_int(nRow), C.c_int(nnz), row.ctypes)
        perm = np.empty(nRow, np.int32)
        length = np.empty(nRow, np.int32)
        jptr = np.empty(K + 1, np.int32)
        jcol = np.empty(nnz, np.int32)
        jval = np.empty(nnz, np.float64)
        pin = None if perm_in is None else _i32(perm_in).ctypes
        ml = self.lib.orc_jds_convert(C.c_int(nRow), C.c_int(nnz), row.ctypes, col.ctypes, val.ctypes,
                                      pin, perm.ctypes, length.ctypes, jptr.ctypes, jcol.ctypes,
                                      jval.ctypes)
        assert ml == K
        return {"maxLength": K, "perm": perm, "length": length, "ptr": jptr, "col_idx": jcol,
                "val": jval}

    def jds_spmv(self, m, x):
        nRow = len(m["perm"])
        y = np.full(nRow, np.nan)
        x = _f64(x)
        self.lib.orc_jds_spmv(C.c_int(nRow), m["perm"].ctypes, m["length"].ctypes, m["ptr"].ctypes,
                              m["col_idx"].ctypes, m["val"].ctypes, x.ctypes, y.ctypes)
        return y

    # ---- DIA (reference src/opt_dia.cpp)
    def dia_convert(self, nRow, nCol, row, col, val):
        row, col, val = _i32(row), _i32(col), _f64(val)
        nnz = len(row)
        nDiag = self.lib.orc_dia_offsets(C.c_int(nRow), C.c_int(nCol), C.c_int(nnz), row.ctypes,
                                         col.ctypes, None)
        ioff = np.empty(nDiag, np.int32)
        self.lib.orc_dia_offsets(C.c_int(nRow), C.c_int(nCol), C.c_int(nnz), row.ctypes, col.ctypes,
                                 ioff.ctypes)
        diag = np.empty((nDiag, nCol), np.float64)
        self.lib.orc_dia_convert(C.c_int(nRow), C.c_int(nCol), C.c_int(nnz), row.ctypes, col.ctypes,
                                 val.ctypes, C.c_int(nDiag), ioff.ctypes, diag.ctypes)
        return {"nDiag": nDiag, "ioff": ioff, "diag": diag, "nRow": nRow, "nCol": nCol}

    def dia_spmv(self, m, x):
        y = np.full(m["nRow"], np.nan)
        x = _f64(x)
        self.lib.orc_dia_spmv(C.c_int(m["nRow"]), C.c_int(m["nCol"]), C.c_int(m["nDiag"]),
                              m["ioff"].ctypes, m["diag"].ctypes, x.ctypes, y.ctypes)
        return y

    # ---- SS (reference src/opt_ss.cpp)
    def ss_convert(self, nRow, row, col, val, W):
        row, col, val = _i32(row), _i32(col), _f64(val)
        nnz = len(row)
        H = self.lib.orc_ss_height(C.c_int(nnz), C.c_int(W))
        row_ptr = np.empty(nRow + 1, np.int32)
        row2d = np.empty(H * W, np.int32)
        col2d = np.empty(H * W, np.int32)
        val2d = np.empty(H * W, np.float64)
        seg = np.empty(H, np.int32)
        nStep = self.lib.orc_ss_convert(C.c_int(nRow), C.c_int(nnz), C.c_int(W), row.ctypes, col.ctypes,
                                        val.ctypes, row_ptr.ctypes, row2d.ctypes, col2d.ctypes,
                                        val2d.ctypes, seg.ctypes)
        counts = np.empty(nStep, np.int32)
        total = self.lib.orc_ss_schedule(C.c_int(H), C.c_int(nStep), seg.ctypes, counts.ctypes, None)
        segs = np.empty(total, np.int32)
        self.lib.orc_ss_schedule(C.c_int(H), C.c_int(nStep), seg.ctypes, counts.ctypes, segs.ctypes)
        return {"W": W, "H": H, "nStep": nStep, "row_ptr": row_ptr, "row_idx": row2d, "col_idx": col2d,
                "val": val2d, "segment_index": seg, "sum_segs_count": counts, "sum_segs": segs,
                "nRow": nRow}

    def ss_spmv(self, m, x, mode="optimized"):
        y = np.full(m["nRow"], np.nan)
        x = _f64(x)
        buf = np.zeros(m["H"] * m["W"], np.float64)
        if mode == "simple":
            self.lib.orc_ss_spmv_simple(C.c_int(m["nRow"]), C.c_int(m["H"]), C.c_int(m["W"]),
                                        m["row_ptr"].ctypes, m["col_idx"].ctypes, m["val"].ctypes,
                                        buf.ctypes, x.ctypes, y.ctypes)
        else:
            self.lib.orc_ss_spmv_optimized(C.c_int(m["nRow"]), C.c_int(m["H"]), C.c_int(m["W"]),
                                           m["row_ptr"].ctypes, m["col_idx"].ctypes, m["val"].ctypes,
                                           C.c_int(m["nStep"]), m["sum_segs_count"].ctypes,
                                           m["sum_segs"].ctypes, buf.ctypes, x.ctypes, y.ctypes)
        return y

    # ---- CSS (reference src/opt_css.cpp)
    def css_convert(self, nRow, nCol, row, col, val, W, nBlockWanted):
        row, col, val = _i32(row), _i32(col), _f64(val)
        nnz = len(row)
        B = self.lib.orc_css_block_width(C.c_int(nCol), C.c_int(nBlockWanted))
        nBlock = self.lib.orc_css_num_blocks(C.c_int(nCol), C.c_int(B))
        Hb = np.empty(nBlock, np.int32)
        bn = np.empty(nBlock, np.int32)
        totalH = self.lib.orc_css_heights(C.c_int(nnz), C.c_int(W), C.c_int(B), C.c_int(nBlock),
                                          col.ctypes, Hb.ctypes, bn.ctypes)
        row_ptr = np.empty(nBlock * (nRow + 1), np.int32)
        row2d = np.empty(totalH * W, np.int32)
        col2d = np.empty(totalH * W, np.int32)
        val2d = np.empty(totalH * W, np.float64)
        seg = np.empty(totalH, np.int32)
        nStep = np.empty(nBlock, np.int32)
        self.lib.orc_css_convert(C.c_int(nRow), C.c_int(nnz), C.c_int(W), C.c_int(B), C.c_int(nBlock),
                                 row.ctypes, col.ctypes, val.ctypes, Hb.ctypes, row_ptr.ctypes,
                                 row2d.ctypes, col2d.ctypes, val2d.ctypes, seg.ctypes, nStep.ctypes)
        counts, segs, s0 = [], [], 0
        for b in range(nBlock):
            c = np.empty(int(nStep[b]), np.int32)
            sb = np.ascontiguousarray(seg[s0:s0 + int(Hb[b])])
            tot = self.lib.orc_css_schedule(C.c_int(int(Hb[b])), C.c_int(int(nStep[b])), sb.ctypes,
                                            c.ctypes, None)
            sg = np.empty(tot, np.int32)
            self.lib.orc_css_schedule(C.c_int(int(Hb[b])), C.c_int(int(nStep[b])), sb.ctypes, c.ctypes,
                                      sg.ctypes)
            counts.append(c)
            segs.append(sg)
            s0 += int(Hb[b])
        cat = lambda xs: np.concatenate(xs).astype(np.int32) if xs else np.empty(0, np.int32)
        return {"W": W, "B": B, "nBlock": nBlock, "totalH": totalH, "H": Hb, "nStep": nStep,
                "row_ptr": row_ptr, "row_idx": row2d, "col_idx": col2d, "val": val2d,
                "segment_index": seg, "sum_segs_count": cat(counts), "sum_segs": cat(segs),
                "nRow": nRow, "nCol": nCol}

    def css_spmv(self, m, x):
        y = np.full(m["nRow"], np.nan)
        x = _f64(x)
        buf = np.zeros(m["totalH"] * m["W"], np.float64)
        cnt = np.ascontiguousarray(m["sum_segs_count"]) if len(m["sum_segs_count"]) else np.zeros(1, np.int32)
        sg = np.ascontiguousarray(m["sum_segs"]) if len(m["sum_segs"]) else np.zeros(1, np.int32)
        self.lib.orc_css_spmv_optimized(C.c_int(m["nRow"]), C.c_int(m["W"]), C.c_int(m["nBlock"]),
                                        m["H"].ctypes, m["row_ptr"].ctypes, m["col_idx"].ctypes,
                                        m["val"].ctypes, m["nStep"].ctypes, cnt.ctypes, sg.ctypes,
                                        buf.ctypes, x.ctypes, y.ctypes)
        return y

    # ---- CSR5 (reference opt/Benchmark_SpMV_using_CSR5, omega = 32; oracle/csr5_oracle.c)
    def csr5_auto_sigma(self, nRow, nnz):
        return int(self.lib.orc_csr5_auto_sigma(C.c_int(nRow), C.c_int(nnz)))

    def csr5_convert(self, nRow, row, col, val, sigma=0):
        row, col, val = _i32(row), _i32(col).copy(), _f64(val).copy()
        nnz = len(row)
        if sigma == 0:
            sigma = self.csr5_auto_sigma(nRow, nnz)
        ptr = np.empty(nRow + 1, np.int32)
        self.lib.orc_crs_row_ptr(C.c_int(nRow), C.c_int(nnz), row.ctypes, ptr.ctypes)
        by, bs, npk, p = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_csr5_shape(C.c_int(nnz), C.c_int(sigma), C.byref(by), C.byref(bs), C.byref(npk), C.byref(p))
        by, bs, npk, p = by.value, bs.value, npk.value, p.value
        tile_ptr = np.zeros(p + 1, np.uint32)
        desc = np.zeros(max(1, p * 32 * npk), np.uint32)
        off_ptr = np.zeros(p + 1, np.int32)
        self.lib.orc_csr5_tile_ptr(C.c_int(nRow), C.c_int(nnz), C.c_int(sigma), C.c_int(p), ptr.ctypes, tile_ptr.ctypes)
        args = (C.c_int(nRow), C.c_int(nnz), C.c_int(sigma), C.c_int(p), C.c_int(by), C.c_int(bs), C.c_int(npk),
                ptr.ctypes, tile_ptr.ctypes, desc.ctypes, off_ptr.ctypes)
        n_off = int(self.lib.orc_csr5_descriptor(*args, None))
        offset = np.zeros(max(1, n_off), np.int32)
        if n_off:
            self.lib.orc_csr5_descriptor(*args, offset.ctypes)
        self.lib.orc_csr5_transpose(C.c_int(nnz), C.c_int(sigma), C.c_int(p), tile_ptr.ctypes, col.ctypes, val.ctypes)
        return {"sigma": sigma, "p": p, "bit_y_offset": by, "bit_scansum_offset": bs, "num_packet": npk,
                "num_offsets": n_off, "row_ptr": ptr, "tile_ptr": tile_ptr, "tile_desc": desc[:p * 32 * npk],
                "tile_desc_offset_ptr": off_ptr, "tile_desc_offset": offset[:n_off], "col_idx": col, "val": val,
                "nRow": nRow}

    def csr5_spmv(self, m, x):
        y = np.full(m["nRow"], np.nan)
        x = _f64(x)
        desc = m["tile_desc"] if len(m["tile_desc"]) else np.zeros(1, np.uint32)
        off = m["tile_desc_offset"] if len(m["tile_desc_offset"]) else np.zeros(1, np.int32)
        self.lib.orc_csr5_spmv(C.c_int(m["nRow"]), C.c_int(len(m["col_idx"])), C.c_int(m["sigma"]), C.c_int(m["p"]),
                               C.c_int(m["bit_y_offset"]), C.c_int(m["bit_scansum_offset"]), C.c_int(m["num_packet"]),
                               m["row_ptr"].ctypes, m["tile_ptr"].ctypes, np.ascontiguousarray(desc).ctypes,
                               m["tile_desc_offset_ptr"].ctypes, np.ascontiguousarray(off).ctypes, m["col_idx"].ctypes,
                               m["val"].ctypes, x.ctypes, y.ctypes)
        return y

    # ---- statistics (reference matrix/script/counter.cpp)
    def counter(self, nRow, nCol, row, col):
        row, col = _i32(row), _i32(col)
        out = np.zeros(5, np.int64)
        self.lib.orc_counter.restype = C.c_double
        var = self.lib.orc_counter(C.c_int(nRow), C.c_int(nCol), C.c_int(len(row)), row.ctypes, col.ctypes, out.ctypes)
        return {"rowMax": int(out[0]), "rowMin": int(out[1]), "colMax": int(out[2]), "colMin": int(out[3]),
                "nDiag": int(out[4]), "rowVar": float(var)}

    # ---- reference verifier and vectors (src/util.cpp:67-102, src/main.cpp:18,31-32)
    def verify(self, nRow, row, col, val, x, y):
        row, col, val, x, y = _i32(row), _i32(col), _f64(val), _f64(x), _f64(y)
        return bool(self.lib.orc_verify(C.c_int(nRow), C.c_int(len(row)), row.ctypes, col.ctypes,
                                        val.ctypes, x.ctypes, y.ctypes))

    def reference_vectors(self, nCol, nRow, seed=3):
        x = np.empty(nCol, np.float64)
        y = np.empty(nRow, np.float64)
        self.lib.orc_reference_vectors(C.c_uint(seed), C.c_int(nCol), C.c_int(nRow), x.ctypes, y.ctypes)
        return x, y

    # ---- synthetic inputs (oracle/synth_oracle.c)
    def stencil(self, kind, n):
        k = {"lap2d5": 0, "lap3d7": 1, "box3d27": 2}[kind]
        nnz = self.lib.synth_stencil_nnz(C.c_int(k), C.c_int(n))
        nRow = self.lib.synth_stencil_rows(C.c_int(k), C.c_int(n))
        row = np.empty(nnz, np.int32)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        got = self.lib.synth_stencil(C.c_int(k), C.c_int(n), row.ctypes, col.ctypes, val.ctypes)
        assert got == nnz
        return int(nRow), int(nRow), row, col, val

    def stencil_rows(self, kind, n, row_begin, row_end):
        k = {"lap2d5": 0, "lap3d7": 1, "box3d27": 2}[kind]
        nRow = int(self.lib.synth_stencil_rows(C.c_int(k), C.c_int(n)))
        cap = (row_end - row_begin) * (5, 7, 27)[k]
        row = np.empty(cap, np.int32)
        col = np.empty(cap, np.int32)
        val = np.empty(cap, np.float64)
        got = self.lib.synth_stencil_range(C.c_int(k), C.c_int(n), C.c_int(row_begin), C.c_int(row_end),
                                           row.ctypes, col.ctypes, val.ctypes)
        return nRow, nRow, row[:got].copy(), col[:got].copy(), val[:got].copy()

    def uniform(self, seed, nRow, nCol, K, row_begin=0, row_end=None):
        row_end = nRow if row_end is None else row_end
        n = (row_end - row_begin) * K
        row = np.empty(n, np.int32)
        col = np.empty(n, np.int32)
        val = np.empty(n, np.float64)
        self.lib.synth_uniform(C.c_uint64(seed), C.c_int(nCol), C.c_int(K), C.c_int(row_begin),
                               C.c_int(row_end), row.ctypes, col.ctypes, val.ctypes)
        return nRow, nCol, row, col, val

    def rmat(self, seed, scale, n_edges):
        row = np.empty(n_edges, np.int32)
        col = np.empty(n_edges, np.int32)
        val = np.empty(n_edges, np.float64)
        nnz = self.lib.synth_rmat(C.c_uint64(seed), C.c_int(scale), C.c_longlong(n_edges), row.ctypes,
                                  col.ctypes, val.ctypes)
        n = 1 << scale
        return n, n, row[:nnz].copy(), col[:nnz].copy(), val[:nnz].copy()


REF_ARRAYS = {
    "crs": {"ptr": np.int32, "idx": np.int32, "val": np.float64},
    "coo": {"row_idx": np.int32, "col_idx": np.int32, "val": np.float64},
    "ell": {"col_idx": np.int32, "val": np.float64},
    "jds": {"perm": np.int32, "length": np.int32, "ptr": np.int32, "col_idx": np.int32,
            "val": np.float64},
    "dia": {"ioff": np.int32, "diag": np.float64},
    "ss": {"row_ptr": np.int32, "row_idx": np.int32, "col_idx": np.int32, "val": np.float64,
           "segment_index": np.int32, "sum_segs_count": np.int32, "sum_segs": np.int32},
    "css": {"H": np.int32, "nStep": np.int32, "col_idx": np.int32, "val": np.float64,
            "row_ptr": np.int32, "sum_segs_count": np.int32, "sum_segs": np.int32},
}
REF_SCALARS = {"crs": [], "coo": [], "ell": ["K"], "jds": ["maxLength"], "dia": ["nDiag"],
               "ss": ["H", "nStep", "W"], "css": ["B", "nBlock", "totalH", "W"]}


def ref_available(variant):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % variant))


class RefPlugin:
    """The reference's own OptimizeProblem/SpMV for one compile-time variant (one use per object;
    the reference keeps its state in globals and never frees)."""

    def __init__(self, variant):
        self.variant = variant
        self.fmt = variant.split("_")[0]
        path = os.path.join(REF_DIR, "libref_%s.so" % variant)
        self.lib = C.CDLL(path)
        self.lib.ref_array.restype = C.c_long
        self.lib.ref_array.argtypes = [C.c_char_p, C.c_void_p]
        self.lib.ref_scalar.argtypes = [C.c_char_p, C.POINTER(C.c_long)]

    def convert(self, nRow, nCol, row, col, val, x):
        row, col, val, x = _i32(row), _i32(col), _f64(val), _f64(x)
        self.nRow = nRow
        self.lib.ref_convert(C.c_int(nRow), C.c_int(nCol), C.c_int(len(row)), row.ctypes, col.ctypes,
                             val.ctypes, x.ctypes)
        out = {}
        for name in ["nRow", "nCol", "nNnz"] + REF_SCALARS[self.fmt]:
            v = C.c_long()
            assert self.lib.ref_scalar(name.encode(), C.byref(v)) == 1, name
            out[name] = int(v.value)
        for name, dt in REF_ARRAYS[self.fmt].items():
            nbytes = self.lib.ref_array(name.encode(), None)
            assert nbytes >= 0, name
            a = np.empty(nbytes // np.dtype(dt).itemsize, dt)
            if nbytes:
                self.lib.ref_array(name.encode(), a.ctypes.data_as(C.c_void_p))
            out[name] = a
        return out

    def spmv(self, x=None):
        if x is not None:
            self.lib.ref_set_x(_f64(x).ctypes)
        y = np.full(self.nRow, np.nan)       # garbage the plugin must fully overwrite
        self.lib.ref_spmv(y.ctypes)
        return y


def ref_csr5_convert(nRow, row, col, val, sigma):
    """The reference's own CSR5 conversion routines (AVX2 twin, omega = 32): oracle/_ref/libref_csr5.so."""
    lib = C.CDLL(os.path.join(REF_DIR, "libref_csr5.so"))
    row, col, val = _i32(row), _i32(col).copy(), _f64(val).copy()
    nnz = len(row)
    ptr = np.searchsorted(row, np.arange(nRow + 1), side="left").astype(np.int32)
    by, bs, npk, p = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.ref_csr5_shape(C.c_int(nnz), C.c_int(sigma), C.byref(by), C.byref(bs), C.byref(npk), C.byref(p))
    by, bs, npk, p = by.value, bs.value, npk.value, p.value
    tile_ptr = np.zeros(p + 1, np.uint32)
    desc = np.zeros(max(1, p * 32 * npk), np.uint32)
    off_ptr = np.zeros(p + 1, np.int32)
    offset = np.zeros(nnz + nRow + 1, np.int32)
    n_off = C.c_int()
    err = lib.ref_csr5_convert(C.c_int(nRow), C.c_int(nnz), C.c_int(sigma), ptr.ctypes, col.ctypes, val.ctypes,
                               tile_ptr.ctypes, desc.ctypes, off_ptr.ctypes, offset.ctypes, C.byref(n_off))
    assert err == 0
    return {"sigma": sigma, "p": p, "bit_y_offset": by, "bit_scansum_offset": bs, "num_packet": npk,
            "num_offsets": n_off.value, "row_ptr": ptr, "tile_ptr": tile_ptr, "tile_desc": desc[:p * 32 * npk],
            "tile_desc_offset_ptr": off_ptr, "tile_desc_offset": offset[:n_off.value], "col_idx": col, "val": val,
            "nRow": nRow}
