# Round 2, GPU call 26 (8 GPUs): final confirmation -- torchrun path (x windows) with the parity leg, single-process path with the
# small-CTA pull kernel, C++ driver over 8 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_8_final.json 2> gpurun_out/r2_bench_peer_8_final.err; echo "bench peer rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_peer_8_final.json")); c = d["config"]
print("N=8 GF %.1f ms %.4f frac %.3f e2e %.2f ms eager %.4f compute-only %.4f exposed %.4f exchange %s timeouts %s launch %s parity %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], c["eager_ms_per_step"], c["compute_only_ms_per_step"],
    c["exposed_exchange_ms"], c["exchange"], c.get("exchange_flag_timeouts"), c["launch"][:14], (d.get("parity") or {}).get("bit_identical")))
PY
timeout 300 python tests/mg_check.py 8 2>&1 | tail -1
B200SPMV_MG_NO_GRAPH=1 timeout 300 python tests/mg_check.py 8 2>&1 | tail -1
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg_dev synth:lap3d7:512 > gpurun_out/driver_r2c26_mg8_dev.txt 2>&1; grep -E "Performance|KernelTime|nGPU|Roofline" gpurun_out/driver_r2c26_mg8_dev.txt
