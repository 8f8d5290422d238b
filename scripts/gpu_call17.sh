source scripts/gpu_check.sh c17
run c3_auto --workload c3 --format auto --steps 10 --no-cpu
run c4_auto --workload c4 --format auto --steps 10 --no-cpu
run c1_crs --workload c1 --format crs --steps 50 --no-cpu
grep -o '"warm_l2[^,]*' gpurun_out/bench_c17_c1_crs.json
run c1_dia --workload c1 --format dia --steps 50 --no-cpu
grep -o '"warm_l2[^,]*' gpurun_out/bench_c17_c1_dia.json
