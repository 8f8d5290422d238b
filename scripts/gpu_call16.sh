source scripts/gpu_check.sh c16
run c5_crs_f32 --workload c5 --format crs --value-f32 --steps 10 --no-cpu
run c1_crs_f32 --workload c1 --format crs --value-f32 --steps 50 --no-cpu
run c3_crs_f32 --workload c3 --format crs --value-f32 --steps 10 --no-cpu
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"; cat gpurun_out/bench_default.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
