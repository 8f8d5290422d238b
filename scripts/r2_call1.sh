# Round 2, GPU call 1: parity suite, mini + full default bench, short-row kernel A/B on c5 / c1.
mkdir -p gpurun_out
TAG=r2c1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/pytest_$TAG.log
run() { # name, args...
  n=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/bench_${TAG}_$n.json 2> gpurun_out/bench_${TAG}_$n.err || echo "bench $n failed rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$n.json"))
    print("$n", d["config"]["format"], "GFLOP/s %.1f ms %.4f frac %.3f e2e %.1f (%.2f ms) conv %.0fms par %s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("convert_ms",0), d.get("parity")), {k: d["config"][k] for k in ("warm_l2_ms_per_step","graph_ms_per_step") if k in d["config"]})
except Exception as e:
    print("$n: no result", e); print(open("gpurun_out/bench_${TAG}_$n.err").read()[-1500:])
PY
}
run mini --mini --steps 5
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2c1_mini.json"))
    for k,v in d.get("configs",{}).items():
        print(k, {f:(round(e.get("gflops",0),1), e.get("parity",{}).get("within_1e-12"), e.get("error")) for f,e in v["formats"].items()}, v.get("cusparse",{}).keys(), v.get("error"))
    print(d.get("gather_ceiling"))
except Exception as e: print("mini configs:", e)
PY
# short-row kernel A/B (c5, c1): row-block stream vs TMA row-chunk stream configurations
B200SPMV_SHORT=rbs run c5_rbs --workload c5 --steps 20 --no-cpu
run c5_tma512s2 --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_S=3 run c5_tma512s3 --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_R=256 run c5_tma256s2 --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_R=256 B200SPMV_TMA_S=3 run c5_tma256s3 --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_R=256 B200SPMV_TMA_S=4 run c5_tma256s4 --workload c5 --steps 20 --no-cpu
B200SPMV_TMA_R=128 B200SPMV_TMA_S=4 run c5_tma128s4 --workload c5 --steps 20 --no-cpu
B200SPMV_SHORT=rbs run c1_rbs --workload c1 --steps 50 --no-cpu
run c1_tma512s2 --workload c1 --steps 50 --no-cpu
B200SPMV_TMA_R=256 run c1_tma256s2 --workload c1 --steps 50 --no-cpu
B200SPMV_TMA_R=128 B200SPMV_TMA_S=3 run c1_tma128s3 --workload c1 --steps 50 --no-cpu
# the default line (all configs), as the driver runs it
run default
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2c1_default.json"))
    for k,v in d.get("configs",{}).items():
        print(k, v.get("error"))
        for f,e in v["formats"].items():
            print("   ", f, "GF %.1f ms %.4f frac %.3f"%(e.get("gflops",0), e.get("ms_per_step",0), e.get("frac",0)), e.get("parity"), e.get("error"), {kk:e[kk] for kk in ("warm_l2_ms_per_step","graph_ms_per_step") if kk in e})
        print("    cusparse", v.get("cusparse"))
    print(d.get("gather_ceiling")); print(d.get("cpu_baseline")); print(d.get("clocks"))
except Exception as e: print("default configs:", e)
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 1200 gpurun_out/bench_${TAG}_reference.json
free -g | head -2; nproc
