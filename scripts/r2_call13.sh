# Round 2, GPU call 13 (8 GPUs): x-window exchange at 8 ranks (dist_check, bench peer + nccl), single-process path with page-locked vectors.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_check.py > gpurun_out/r2_dist_check_8.log 2>&1; echo "dist_check rc=$?"; grep dist_check gpurun_out/r2_dist_check_8.log | cut -c1-200
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    c = d["config"]
    print(sys.argv[1], "GF %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms, cpus %s) eager %.4f compute-only %.4f exposed %.4f exchange %s (%s) launch %s parity %s" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_cpus_bound"), c.get("eager_ms_per_step", 0),
        c.get("compute_only_ms_per_step", 0), c.get("exposed_exchange_ms", 0), c.get("exchange"), c.get("exchange_fallback_reason"), c.get("launch", "")[:20], (d.get("parity") or {}).get("bit_identical")))
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
}
timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_8.json 2> gpurun_out/r2_bench_peer_8.err; echo "bench peer rc=$?"; show gpurun_out/r2_bench_peer_8.json; tail -2 gpurun_out/r2_bench_peer_8.err | cut -c1-300
B200SPMV_DIST_EXCHANGE=nccl timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_nccl_8.json 2> gpurun_out/r2_bench_nccl_8.err; echo "bench nccl rc=$?"; show gpurun_out/r2_bench_nccl_8.json
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR4 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_peer_4.json 2> gpurun_out/r2_bench_peer_4.err; echo "bench peer 4 rc=$?"; show gpurun_out/r2_bench_peer_4.json
timeout 300 python tests/mg_check.py 8 2>&1 | tail -1
