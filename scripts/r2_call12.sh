# Round 2, GPU call 12 (1 GPU): COO entry stream variants (bit 0: CTA barrier per tile, bit 1: branch-free lane sums).
mkdir -p gpurun_out
TAG=r2c12
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
for v in 0 1 2 3; do b c5_coo_v$v B200SPMV_COO_VARIANT=$v; done
for v in 1 3; do b c5_coo_e2048_v$v B200SPMV_COO_VARIANT=$v B200SPMV_COO_E=2048; done
b c5_coo_v3_c7 B200SPMV_COO_VARIANT=3 B200SPMV_COO_CTAS=7
BARGS="--workload c3 --format coo"
for v in 1 3; do b c3_coo_v$v B200SPMV_COO_VARIANT=$v; done
for v in 1 3; do b c3_coo_e2048_v$v B200SPMV_COO_VARIANT=$v B200SPMV_COO_E=2048; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -2 gpurun_out/pytest_$TAG.log
