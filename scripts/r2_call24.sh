# Round 2, GPU call 24 (1 GPU): two-phase CRS entry stream (all gathers of a chunk issued before the reductions).
mkdir -p gpurun_out
TAG=r2c24
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
BARGS="--workload c4 --format crs"; b c4_crs_es B200SPMV_CRS_PATH=es; b c4_crs_es_c4 B200SPMV_CRS_PATH=es B200SPMV_ES_CTAS=4; b c4_crs_es_1024 B200SPMV_CRS_PATH=es B200SPMV_ES_E=1024
BARGS="--workload c3 --format crs"; b c3_crs_es X=1
BARGS="--workload c2 --format crs"; b c2_crs_es X=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "entry_stream or guard_bands" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -2 gpurun_out/pytest_$TAG.log
