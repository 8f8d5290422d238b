# Round 2, GPU call 21 (1 GPU): entry streams with the adaptive scan depth.
mkdir -p gpurun_out
TAG=r2c21
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
BARGS="--workload c5 --format coo"; b c5_coo X=1
BARGS="--workload c4 --format coo"; b c4_coo X=1
BARGS="--workload c3 --format coo"; b c3_coo X=1
BARGS="--workload c4 --format crs"; b c4_crs_es X=1; b c4_crs_es_e2048 B200SPMV_ES_E=2048; b c4_crs_es_c8 B200SPMV_ES_CTAS=8
BARGS="--workload c3 --format crs"; b c3_crs_es X=1
BARGS="--workload c5 --format crs --format crs"; b c5_crs X=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -2 gpurun_out/pytest_$TAG.log
