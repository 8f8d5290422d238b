# Round 2, GPU call 7 (8 GPUs): torchrun path (dist_check, bench at 8) and the single-process path (mg_check 8, 4).
bash scripts/r2_multi.sh 8 quick
timeout 400 python tests/mg_check.py 8 2>&1 | tail -2
timeout 300 python tests/mg_check.py 4 2>&1 | tail -1
# global nnz beyond int32 (3.58 G), every block below it: the multi-GPU paths lift the reference's 2^31 ceiling (src/util.h:8)
MG_CHECK_P0=800 MG_CHECK_NO_SINGLE=1 timeout 600 python tests/mg_check.py 8 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_multi_4.json 2> gpurun_out/r2_bench_multi_4.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_multi_4.json")); c = d["config"]
    print("N=4 GF %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms) compute-only %.4f exposed %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], c.get("compute_only_ms_per_step", 0), c.get("exposed_exchange_ms", 0)))
except Exception as e:
    print("N=4 no result", e)
PY
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg_dev synth:lap3d7:320 > gpurun_out/driver_r2c7_mg8_dev.txt 2>&1; grep -E "Performance|KernelTime|nGPU|Halo|Graph|invalid|failed" gpurun_out/driver_r2c7_mg8_dev.txt
