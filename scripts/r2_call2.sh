# Round 2, GPU call 2: parity suite with the generalized row-chunk stream + column-block engines; c2 / c4 A/B; e2e knobs.
mkdir -p gpurun_out
TAG=r2c2
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
# c2: column-block layouts
B200SPMV_COL_BLOCKS=0 run c2_ell_plain --workload c2 --format ell --steps 10
run c2_ell_crs3 --workload c2 --format ell --steps 10
B200SPMV_COLBLOCK=cbs run c2_ell_cbs3 --workload c2 --format ell --steps 10
B200SPMV_COL_BLOCKS=2 run c2_ell_crs2 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_COL_BLOCKS=4 run c2_ell_crs4 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_COL_BLOCKS=4 B200SPMV_COLBLOCK=cbs run c2_ell_cbs4 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TMA_RL=256 run c2_ell_crs3_r256 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TMA_RL=256 B200SPMV_TMA_S=3 run c2_ell_crs3_r256s3 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_TMA_S=3 run c2_ell_crs3_s3 --workload c2 --format ell --steps 10 --no-cpu
run c2_jds --workload c2 --format jds --steps 10
run c2_ss --workload c2 --format ss --steps 10
run c2_css3 --workload c2 --format css --n-block 3 --steps 10
B200SPMV_TMA_MAXLEN=16 run c2_css3_tile --workload c2 --format css --n-block 3 --steps 10 --no-cpu
# c4 CRS: looped row-chunk stream vs tile-stream
run c4_crs --workload c4 --format crs --steps 20
B200SPMV_TMA_MAXLEN=16 run c4_crs_tile --workload c4 --format crs --steps 20 --no-cpu
B200SPMV_TMA_RL=256 run c4_crs_r256 --workload c4 --format crs --steps 20 --no-cpu
B200SPMV_TMA_S=3 run c4_crs_s3 --workload c4 --format crs --steps 20 --no-cpu
run c4_ss --workload c4 --format ss --steps 20 --no-cpu
# c3: CSS as column-blocked CRS
run c3_css2 --workload c3 --format css --n-block 2 --steps 10 --no-cpu
run c3_css3 --workload c3 --format css --n-block 3 --steps 10 --no-cpu
# c5 / c1 with the generalized kernel (regression check against call 1: 920 / 21 us)
run c5_crs --workload c5 --steps 20 --no-cpu
run c1_crs --workload c1 --steps 50 --no-cpu
# e2e knobs on c5
B200SPMV_HOST_CHUNKS=32 run c5_e2e_c32 --workload c5 --steps 10 --no-cpu
B200SPMV_HOST_CHUNKS=8 run c5_e2e_c8 --workload c5 --steps 10 --no-cpu
B200SPMV_HOST_CHUNKS=4 run c5_e2e_c4 --workload c5 --steps 10 --no-cpu
B200SPMV_HOST_CHUNKS=1 B200SPMV_HOST_PIECES=1 run c5_e2e_c1 --workload c5 --steps 10 --no-cpu
# plugin binary, host semantics, c1 (the reference driver's own path)
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=5 singlespmv_b200/plugin/bin/spmv_b200_crs synth:lap2d5:1024 > gpurun_out/driver_${TAG}_c1_crs_host.txt 2>&1; grep -E "Performance|KernelTime|VectorRes" gpurun_out/driver_${TAG}_c1_crs_host.txt
B200_NO_HOST_REGISTER=1 SPMV_MIN_SECONDS=0.3 SPMV_NTRY=5 singlespmv_b200/plugin/bin/spmv_b200_crs synth:lap2d5:1024 > gpurun_out/driver_${TAG}_c1_crs_host_pageable.txt 2>&1; grep -E "Performance|KernelTime" gpurun_out/driver_${TAG}_c1_crs_host_pageable.txt
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 singlespmv_b200/plugin/bin/spmv_b200_ss_prof synth:lap2d5:1024 > gpurun_out/driver_${TAG}_c1_ss_prof.txt 2>&1; grep -E "nStep|StepCount|MulPerf|SumPerf|Performance" gpurun_out/driver_${TAG}_c1_ss_prof.txt | head
