# Round 2, GPU call 10 (1 GPU): row-blocked tests, COO defaults, ncu capture of coo_stream_kernel on c5.
mkdir -p gpurun_out
TAG=r2c10
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo or row_blocked or int32_entries" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
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
b c5_coo X=1
b c5_coo_c7 B200SPMV_COO_CTAS=7
b c5_coo_c5 B200SPMV_COO_CTAS=5
b c5_coo_e2048 B200SPMV_COO_E=2048
BARGS="--workload c3 --format coo"
b c3_coo X=1
b c3_coo_e2048 B200SPMV_COO_E=2048
BARGS="--workload c4 --format coo"
b c4_coo X=1
BARGS="--workload c1 --format coo"
b c1_coo X=1
A="--steps 3 --warmup 3 --no-cpu"
python bench.py --workload c5 --format coo $A > gpurun_out/plain_c5_coo.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:coo_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_c5_coo python bench.py --workload c5 --format coo $A > gpurun_out/ncu_c5_coo.log 2>&1
echo "ncu rc=$?"
