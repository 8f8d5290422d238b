"""Generate tests/golden/*.npz from the reference's OWN code (run in the build container only).

Inputs : the reference's fixtures /root/reference/matrix/test/{3x3,5x5,10x10,random}.mtx, read
         with the loader semantics of /root/reference/src/util.cpp:30-66 (first non-'%' line is
         the header, exactly L triples, 1-based -> 0-based, sort by (row, col)), plus mini
         versions of the five BASELINE.json shapes from oracle/synth_oracle.c.
Vectors: x = glibc rand() stream seeded with 3, x before y (src/main.cpp:18,31-32).
Outputs: every SpMatOpt array and the SpMV result y of every reference plugin variant built by
         oracle/Makefile (oracle/_ref/libref_*.so = unmodified /root/reference/src/opt_*.cpp), and the CSR5
         arrays (sigma 4 and 16) from the vendored conversion routines (oracle/_ref/libref_csr5.so).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, RefPlugin, build_oracle, ref_csr5_convert  # noqa: E402

VARIANTS = ["crs", "coo", "ell", "jds", "dia", "ss_simple_w4", "ss_opt_w2", "ss_opt_w4", "ss_opt_w32",
            "css_simple_w4_n2", "css_opt_w2_n2", "css_opt_w4_n3", "css_opt_w32_n4"]


def load_like_reference(path):
    lines = open(path).read().split("\n")
    i = 0
    while lines[i].startswith("%"):
        i += 1
    M, N, L = (int(t) for t in lines[i].split()[:3])
    toks = " ".join(lines[i + 1:]).split()
    ent = []
    for k in range(L):
        r, c, v = int(toks[3 * k]) - 1, int(toks[3 * k + 1]) - 1, float(toks[3 * k + 2])
        ent.append((r, c, v))
    ent.sort(key=lambda e: (e[0], e[1]))
    row = np.array([e[0] for e in ent], np.int32)
    col = np.array([e[1] for e in ent], np.int32)
    val = np.array([e[2] for e in ent], np.float64)
    return M, N, row, col, val


def main():
    build_oracle()
    orc = Oracle()
    cases = {}
    for name in ["3x3", "5x5", "10x10", "random"]:
        cases["fixture_" + name] = load_like_reference("/root/reference/matrix/test/%s.mtx" % name)
    cases["mini_lap2d5_n6"] = orc.stencil("lap2d5", 6)
    cases["mini_lap3d7_n4"] = orc.stencil("lap3d7", 4)
    cases["mini_box3d27_n3"] = orc.stencil("box3d27", 3)
    cases["mini_uniform_48x8"] = orc.uniform(1, 48, 48, 8)
    cases["mini_rmat_s6"] = orc.rmat(42, 6, 700)
    cases["mini_rmat_s9"] = orc.rmat(42, 9, 6000)      # empty rows + long rows over many CSR5 tiles
    for name, (nRow, nCol, row, col, val) in cases.items():
        key = row.astype(np.int64) * nCol + col
        assert np.all(np.diff(key) > 0), name + ": not sorted / has duplicates"
        x, _ = orc.reference_vectors(nCol, nRow, 3)
        out = {"nRow": np.int64(nRow), "nCol": np.int64(nCol), "in_row": row, "in_col": col,
               "in_val": val, "x": x}
        for v in VARIANTS:
            p = RefPlugin(v)
            m = p.convert(nRow, nCol, row, col, val, x)
            y1 = p.spmv()
            y2 = p.spmv()                      # the reference verifies twice (src/main.cpp:40-56)
            if v == "coo":                   # omp atomic scatter: summation order is not fixed
                assert np.allclose(y1, y2, rtol=1e-13, atol=0), (name, v)
            else:
                assert np.array_equal(y1, y2), (name, v)
            assert orc.verify(nRow, row, col, val, x, y1), (name, v)
            for k, a in m.items():
                out["%s.%s" % (v, k)] = np.asarray(a)
            out["%s.y" % v] = y1
        # CSR5 arrays from the reference's conversion routines (AVX2 twin at omega = 32, oracle/ref_csr5_shim.cpp)
        for sigma in (4, 16):
            m = ref_csr5_convert(nRow, row, col, val, sigma)
            for k, a in m.items():
                if k not in ("nRow", "row_ptr"):
                    out["csr5_s%d.%s" % (sigma, k)] = np.asarray(a)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, nRow, nCol, len(row), "ok")


if __name__ == "__main__":
    main()
