# Final single-GPU regression of the round: full GPU test suite, smoke, default bench (+ reference arm), key configs.
source scripts/gpu_check.sh final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"; cut -c1-700 gpurun_out/bench_default.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json
run c5_crs --workload c5 --format crs --steps 10 --no-cpu
run c1_crs --workload c1 --format crs --steps 50 --no-cpu
run c4_dia --workload c4 --steps 20 --no-cpu
run c3_auto --workload c3 --format auto --steps 10 --no-cpu --compare-cusparse
