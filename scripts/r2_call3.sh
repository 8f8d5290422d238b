# Round 2, GPU call 3: parity suite; tile-stream LDG vs TMA on the gather-bound configs; column-block engine; HYB.
mkdir -p gpurun_out
TAG=r2c3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/pytest_$TAG.log
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/bench_${TAG}_$n.json 2> gpurun_out/bench_${TAG}_$n.err || echo "bench $n failed rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$n.json"))
    print("$n", d["config"]["format"], "GFLOP/s %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms) conv %.0fms par %s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("convert_ms",0), (d.get("parity") or {}).get("bit_identical")), {k: d["config"][k] for k in ("warm_l2_ms_per_step","graph_ms_per_step") if k in d["config"]})
except Exception as e:
    print("$n: no result", e); print(open("gpurun_out/bench_${TAG}_$n.err").read()[-1500:])
PY
}
run c2_ell --workload c2 --format ell --steps 10
B200SPMV_TS_LOAD=tma run c2_ell_tma --workload c2 --format ell --steps 10
B200SPMV_TS_LOAD=tma B200SPMV_COL_BLOCKS=4 run c2_ell_tma4 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma B200SPMV_COL_BLOCKS=2 run c2_ell_tma2 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma B200SPMV_TS_THREADS=128 run c2_ell_tma_t128 --workload c2 --format ell --steps 10 --no-cpu
run c2_css3 --workload c2 --format css --n-block 3 --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma run c2_css3_tma --workload c2 --format css --n-block 3 --steps 10 --no-cpu
run c2_jds --workload c2 --format jds --steps 10 --no-cpu
run c2_ss --workload c2 --format ss --steps 10 --no-cpu
run c3_crs --workload c3 --format crs --steps 10 --no-cpu
B200SPMV_TS_LOAD=tma run c3_crs_tma --workload c3 --format crs --steps 10
run c4_crs --workload c4 --format crs --steps 20 --no-cpu
B200SPMV_TS_LOAD=tma run c4_crs_tma --workload c4 --format crs --steps 20
run c4_hyb --workload c4 --format hyb --steps 20
run c3_hyb --workload c3 --format hyb --steps 10
run c5_crs --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_R=512 run c5_crs_512 --workload c5 --steps 20 --no-cpu
run c5_css --workload c5 --format css --steps 20 --no-cpu
run c1_crs --workload c1 --steps 50 --no-cpu
