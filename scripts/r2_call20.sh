# Round 2, GPU call 20 (1 GPU): CRS entry stream -- correctness on the goldens/skew cases (tolerance), c2..c5 against the tile-stream.
mkdir -p gpurun_out
TAG=r2c20
python - <<'PY'
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import singlespmv_b200 as sp
from oracle_lib import Oracle
from conftest import skewed_matrix
o = Oracle()
rng = np.random.default_rng(7)
cases = []
for nRow, nCol, d in [(257, 301, 12), (1000, 777, 40), (5000, 5000, 6), (33, 4000, 3), (20000, 20000, 30)]:
    row, col, val = skewed_matrix(rng, nRow, nCol, d)
    cases.append((nRow, nCol, row, col, val))
for k in ("lap2d5", "box3d27"):
    nr, nc, r, c, v = o.stencil(k, 23)
    cases.append((nr, nc, r, c, v))
nr, nc, r, c, v = o.rmat(42, 13, 400000); cases.append((nr, nc, r, c, v))
n = 9000
cases.append((3, n, np.full(n, 1, np.int32), np.arange(n, dtype=np.int32), rng.standard_normal(n)))
bad = 0
for nRow, nCol, row, col, val in cases:
    x = rng.random(nCol)
    y_ref = o.crs_result(nRow, row, col, val, x)
    mag = np.zeros(nRow); np.add.at(mag, row, np.abs(val * x[col]))
    A = sp.SpMat(nRow, nCol, row, col, val)
    for path in (4,):
        A_opt, x_opt = sp.OptimizeProblem(A, sp.Vec(x), "crs", crs_path=path)
        y = sp.Vec(np.full(nRow, np.nan)); sp.SpMV(A_opt, x_opt, y)
        err = np.abs(y.val - y_ref)
        okr = (err <= 1e-12 * np.abs(y_ref)) | (err <= 1e-12 * mag)
        print("rows %6d nnz %8d crs_kernel %d: rows off %d, max err/mag %.2e, finite %s" % (nRow, len(row), A_opt.scalar("crs_kernel"), int((~okr).sum()),
              float(np.max(err / np.maximum(mag, 1e-300))) if nRow else 0.0, bool(np.all(np.isfinite(y.val)))), flush=True)
        bad += int((~okr).sum())
print("ENTRY STREAM PARITY", "OK" if bad == 0 else "FAILED")
PY
b() { # tag, env..., -- bench args
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 $BARGS > gpurun_out/bench_${TAG}_$tag.json 2> gpurun_out/bench_${TAG}_$tag.err
  python - gpurun_out/bench_${TAG}_$tag.json $tag <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "GF %.1f ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
for w in c3 c4 c2; do
BARGS="--workload $w --format crs"
b ${w}_crs_es X=1
b ${w}_crs_tile B200SPMV_CRS_PATH=tile
done
BARGS="--workload c4 --format crs"
b c4_crs_es_e2048 B200SPMV_ES_E=2048
BARGS="--workload c3 --format crs"
b c3_crs_es_e1024 B200SPMV_ES_E=1024
