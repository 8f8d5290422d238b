# Round 2, GPU call 5 (2 GPUs): full parity suite incl. fp32 / mg / HYB, single-process multi-GPU check, TMA policy A/B.
mkdir -p gpurun_out
TAG=r2c5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python tests/mg_check.py 2 2>&1 | tail -3
B200SPMV_MG_NO_GRAPH=1 timeout 300 python tests/mg_check.py 2 2>&1 | tail -2
timeout 300 python tests/mg_check.py 1 2>&1 | tail -2
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 B200_NGPU=2 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg_dev synth:lap3d7:256 > gpurun_out/driver_${TAG}_mg2_dev.txt 2>&1; grep -E "Performance|KernelTime|nGPU|Halo|Graph|invalid" gpurun_out/driver_${TAG}_mg2_dev.txt
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 B200_NGPU=2 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg synth:lap3d7:256 > gpurun_out/driver_${TAG}_mg2_host.txt 2>&1; grep -E "Performance|KernelTime|nGPU|invalid" gpurun_out/driver_${TAG}_mg2_host.txt
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/bench_${TAG}_$n.json 2> gpurun_out/bench_${TAG}_$n.err || echo "bench $n failed rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$n.json"))
    print("$n", d["config"]["format"], "GFLOP/s %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms) par %s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity")))
except Exception as e:
    print("$n: no result", e); print(open("gpurun_out/bench_${TAG}_$n.err").read()[-1500:])
PY
}
B200SPMV_TS_LOAD=tma B200SPMV_TMA_POLICY=1 run c2_ell_tma_p1 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma B200SPMV_TMA_POLICY=2 run c2_ell_tma_p2 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma B200SPMV_TMA_POLICY=3 run c2_ell_tma_p3 --workload c2 --format ell --steps 10 --no-cpu
run c5_crs_f32 --workload c5 --format crs --precision 1 --steps 20
run c5_ell_f32 --workload c5 --format ell --precision 1 --steps 20
run c2_ell_f32 --workload c2 --format ell --precision 1 --steps 10
run c3_crs_f32 --workload c3 --format crs --precision 1 --steps 10
