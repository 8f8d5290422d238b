# Round 2 multi-GPU check + bench.  Usage: bash scripts/r2_multi.sh N [quick]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tests/dist_check.py > gpurun_out/r2_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep dist_check gpurun_out/r2_dist_check_$N.log; tail -5 gpurun_out/r2_dist_check_$N.log | grep -v dist_check | tail -3
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    c = d["config"]
    print(sys.argv[1], "GF %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms) eager %.4f compute-only %.4f exposed %.4f kernel %s parity %s" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], c.get("eager_ms_per_step", 0),
        c.get("compute_only_ms_per_step", 0), c.get("exposed_exchange_ms", 0), c.get("local_kernel"), (d.get("parity") or {}).get("bit_identical")))
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
}
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_multi_$N.json 2> gpurun_out/r2_bench_multi_$N.err; echo "bench rc=$?"; show gpurun_out/r2_bench_multi_$N.json; tail -2 gpurun_out/r2_bench_multi_$N.err
if [ "$2" != "quick" ]; then
B200SPMV_DIST_CRS_PATH=1 timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_multi_${N}_tile.json 2> gpurun_out/r2_bench_multi_${N}_tile.err; echo "bench tile rc=$?"; show gpurun_out/r2_bench_multi_${N}_tile.json
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-graph > gpurun_out/r2_bench_multi_${N}_eager.json 2> gpurun_out/r2_bench_multi_${N}_eager.err; echo "bench eager rc=$?"; show gpurun_out/r2_bench_multi_${N}_eager.json
B200SPMV_TMA_CTAS=4 timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_multi_${N}_ctas4.json 2> gpurun_out/r2_bench_multi_${N}_ctas4.err; echo "bench ctas4 rc=$?"; show gpurun_out/r2_bench_multi_${N}_ctas4.json
fi
timeout 300 python bench.py --impl reference --gpus $N --steps 5 --warmup 1 --sample > gpurun_out/r2_bench_ref_$N.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_ref_$N.json
