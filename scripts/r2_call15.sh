# Round 2, GPU call 15 (1 GPU): COO entry stream fed by LDG instead of TMA (gather-bound matrices).
mkdir -p gpurun_out
TAG=r2c15
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
BARGS="--workload c3 --format coo"
b c3_coo_tma X=1
b c3_coo_ldg B200SPMV_COO_PATH=ldg
b c3_coo_ldg_e2048 B200SPMV_COO_PATH=ldg B200SPMV_COO_E=2048
BARGS="--workload c5 --format coo"
b c5_coo_tma X=1
b c5_coo_ldg B200SPMV_COO_PATH=ldg
b c5_coo_ldg_e2048 B200SPMV_COO_PATH=ldg B200SPMV_COO_E=2048
BARGS="--workload c2 --format coo"
b c2_coo_tma X=1
b c2_coo_ldg B200SPMV_COO_PATH=ldg
BARGS="--workload c4 --format coo"
b c4_coo_ldg B200SPMV_COO_PATH=ldg
B200SPMV_COO_PATH=ldg timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -2 gpurun_out/pytest_$TAG.log
