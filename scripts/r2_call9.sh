# Round 2, GPU call 9 (1 GPU): COO entry stream with in-CTA stitching + prefetched tile-front row; launch list.
mkdir -p gpurun_out
TAG=r2c9
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo or hyb or row_blocked or int32_entries" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
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
BARGS="--workload c5 --format coo"
b c5_coo_e2048 X=1
b c5_coo_e1024 B200SPMV_COO_E=1024
b c5_coo_e1024_c6 B200SPMV_COO_E=1024 B200SPMV_COO_CTAS=6
BARGS="--workload c3 --format coo"
b c3_coo_e1024 B200SPMV_COO_E=1024
BARGS="--workload c4 --format coo"
b c4_coo_e1024 B200SPMV_COO_E=1024
BARGS="--workload c1 --format coo"
b c1_coo_e1024 B200SPMV_COO_E=1024
B200SPMV_COO_E=1024 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c5_coo.csv python bench.py --workload c5 --format coo --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches_c5_coo.log 2>&1; echo "launch list rc=$?"
grep -E "coo_stream" gpurun_out/r2_launches_c5_coo.csv | tail -4
