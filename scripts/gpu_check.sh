# One-GPU check: parity tests, bench lines, optional ncu.  Usage: bash scripts/gpu_check.sh [tag]
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
run() { # name, args...
  n=$1; shift
  timeout 300 python bench.py "$@" > gpurun_out/bench_${TAG}_$n.json 2> gpurun_out/bench_${TAG}_$n.err || echo "bench $n failed rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$n.json"))
    print("$n", d["config"]["format"], "GFLOP/s %.1f ms %.4f frac %.3f e2e %.1f conv %.0fms"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["config"].get("convert_ms",0)))
except Exception as e:
    print("$n: no result", e); print(open("gpurun_out/bench_${TAG}_$n.err").read()[-1500:])
PY
}
