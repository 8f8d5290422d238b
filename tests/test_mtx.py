"""Matrix-Market ingest (b200spmv_load_mtx): banner-aware mode against scipy.io.mmread, reference mode against the
reference loader's semantics (src/util.cpp:30-66, restated in tests/golden/make_golden.py::load_like_reference and
pinned by the goldens' in_row/in_col/in_val)."""
import os

import numpy as np
import pytest

from conftest import load_golden


def write(path, banner, M, N, entries, comments=("% a comment",)):
    with open(path, "w") as f:
        if banner:
            f.write(banner + "\n")
        for c in comments:
            f.write(c + "\n")
        f.write("%d %d %d\n" % (M, N, len(entries)))
        for e in entries:
            f.write(" ".join(repr(v) if isinstance(v, float) else str(v) for v in e) + "\n")


def test_load_mtx_errors_on_cpu(tmp_path):
    import singlespmv_b200 as sp
    with pytest.raises(sp.B200SpmvError) as e:
        sp.DeviceCoo.from_mtx(tmp_path / "missing.mtx")
    assert e.value.status == -1
    p = tmp_path / "cplx.mtx"
    write(p, "%%MatrixMarket matrix coordinate complex general", 2, 2, [(1, 1, 1.0, 0.0)])
    with pytest.raises(sp.B200SpmvError) as e:
        sp.DeviceCoo.from_mtx(p)
    assert e.value.status == -3
    p = tmp_path / "oob.mtx"
    write(p, "%%MatrixMarket matrix coordinate real general", 2, 2, [(3, 1, 1.0)])
    with pytest.raises(sp.B200SpmvError):
        sp.DeviceCoo.from_mtx(p)
    p = tmp_path / "short.mtx"
    write(p, None, 3, 3, [(1, 1, 1.0)])
    open(p, "a").close()
    txt = open(p).read().replace("3 3 1", "3 3 4")
    open(p, "w").write(txt)
    with pytest.raises(sp.B200SpmvError):
        sp.DeviceCoo.from_mtx(p, reference_semantics=True)
    if sp.device_count() == 0:                     # parsed fine, but there is nowhere to put it
        p = tmp_path / "ok.mtx"
        write(p, "%%MatrixMarket matrix coordinate real general", 2, 2, [(1, 1, 1.0)])
        with pytest.raises(sp.B200SpmvError) as e:
            sp.DeviceCoo.from_mtx(p)
        assert e.value.status == -2 and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("field,symmetry", [("real", "general"), ("real", "symmetric"), ("pattern", "general"),
                                            ("pattern", "symmetric"), ("integer", "general"), ("real", "skew-symmetric")])
def test_load_mtx_banner_vs_scipy(tmp_path, field, symmetry):
    import scipy.io
    import singlespmv_b200 as sp
    rng = np.random.default_rng(hash((field, symmetry)) % 1000)
    n = 300
    r = rng.integers(0, n, 4000)
    c = rng.integers(0, n, 4000)
    if symmetry != "general":
        keep = r > c if symmetry == "skew-symmetric" else r >= c
        r, c = r[keep], c[keep]
    key = np.unique(r.astype(np.int64) * n + c)        # Matrix-Market files list a coordinate once
    r, c = (key // n).astype(int), (key % n).astype(int)
    if field == "real":
        ent = [(int(a) + 1, int(b) + 1, float(v)) for a, b, v in zip(r, c, rng.standard_normal(len(r)))]
    elif field == "integer":
        ent = [(int(a) + 1, int(b) + 1, int(v)) for a, b, v in zip(r, c, rng.integers(-9, 10, len(r)))]
    else:
        ent = [(int(a) + 1, int(b) + 1) for a, b in zip(r, c)]
    order = rng.permutation(len(ent))
    p = tmp_path / "m.mtx"
    write(p, "%%%%MatrixMarket matrix coordinate %s %s" % (field, symmetry), n, n, [ent[i] for i in order],
          comments=("% generated", "%", "% more"))
    d = sp.DeviceCoo.from_mtx(p)
    nRow, nCol, row, col, val = d.to_host()
    ref = scipy.io.mmread(str(p)).tocsr()
    ref.sort_indices()
    ref = ref.tocoo()
    assert (nRow, nCol) == ref.shape
    assert np.array_equal(row, ref.row) and np.array_equal(col, ref.col) and np.array_equal(val, ref.data.astype(np.float64))
    # the result satisfies the plugins' input contract: convert + multiply
    x = rng.random(n)
    A_opt, x_opt = sp.OptimizeProblem(sp.SpMat(nRow, nCol, row, col, val), sp.Vec(x), "crs")
    B = sp.SpMatOpt("crs").convert_device(d)
    y1, y2 = np.empty(n), np.empty(n)
    A_opt.multiply_host(x, y1)
    B.multiply_host(x, y2)
    assert np.array_equal(y1, y2) and np.allclose(y1, ref.tocsr() @ x, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_load_mtx_sums_duplicates(tmp_path):
    import singlespmv_b200 as sp
    p = tmp_path / "dup.mtx"
    write(p, "%%MatrixMarket matrix coordinate real general", 3, 4, [(2, 3, 1.0), (1, 1, 2.0), (2, 3, 4.0), (3, 4, 8.0), (2, 3, 16.0)])
    _, _, row, col, val = sp.DeviceCoo.from_mtx(p).to_host()
    assert row.tolist() == [0, 1, 2] and col.tolist() == [0, 2, 3] and val.tolist() == [2.0, 21.0, 8.0]
    # reference semantics keep them (src/util.cpp:44-51) -- and the plugins then refuse the matrix
    d = sp.DeviceCoo.from_mtx(p, reference_semantics=True)
    _, _, row, col, val = d.to_host()
    assert row.tolist() == [0, 1, 1, 1, 2] and col.tolist() == [0, 2, 2, 2, 3] and sorted(val[1:4].tolist()) == [1.0, 4.0, 16.0]
    with pytest.raises(sp.B200SpmvError):
        sp.SpMatOpt("crs").convert_device(d)


@pytest.mark.gpu
def test_load_mtx_duplicates_sum_in_file_order(tmp_path):
    """duplicates are added left to right in file order (a non-associative fp64 sum), exactly like a host
    loop over the file -- not in whatever order a library reduction picks."""
    import singlespmv_b200 as sp
    rng = np.random.default_rng(11)
    ents = []
    for k in range(4000):
        r, c = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        ents.append((r, c, float(rng.standard_normal() * 10.0 ** int(rng.integers(-8, 8)))))
    p = tmp_path / "dups.mtx"
    write(p, "%%MatrixMarket matrix coordinate real general", 40, 40, ents)
    _, _, row, col, val = sp.DeviceCoo.from_mtx(p).to_host()
    text = {}
    for ln in open(p).read().splitlines()[3:]:               # banner, comment, size line
        r, c, v = ln.split()
        k = (int(r) - 1, int(c) - 1)
        text[k] = text[k] + float(v) if k in text else float(v)
    assert len(row) == len(text)
    for r, c, v in zip(row, col, val):
        assert v == text[(int(r), int(c))], (r, c)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fixture_10x10", "fixture_random", "fixture_5x5", "mini_rmat_s9"])
def test_load_mtx_reference_semantics(tmp_path, name):
    """Same triples the reference's loader produced for its own fixtures (goldens), whatever the banner says."""
    import singlespmv_b200 as sp
    g = load_golden(name)
    rng = np.random.default_rng(2)
    order = rng.permutation(len(g["in_row"]))
    ent = [(int(g["in_row"][i]) + 1, int(g["in_col"][i]) + 1, float(g["in_val"][i])) for i in order]
    p = tmp_path / (name + ".mtx")
    # a 'symmetric' banner must be ignored in this mode, exactly like src/util.cpp does
    write(p, "%%MatrixMarket matrix coordinate real symmetric", int(g["nRow"]), int(g["nCol"]), ent)
    nRow, nCol, row, col, val = sp.DeviceCoo.from_mtx(p, reference_semantics=True).to_host()
    assert (nRow, nCol) == (int(g["nRow"]), int(g["nCol"]))
    assert np.array_equal(row, g["in_row"]) and np.array_equal(col, g["in_col"]) and np.array_equal(val, g["in_val"])
