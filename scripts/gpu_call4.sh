source scripts/gpu_check.sh c4
run c2_ss --workload c2 --format ss --steps 10 --no-cpu
run c2_css4 --workload c2 --format css --n-block 4 --steps 10 --no-cpu
run c2_css8 --workload c2 --format css --n-block 8 --steps 10 --no-cpu
run c2_css16 --workload c2 --format css --n-block 16 --steps 10 --no-cpu
run c2_coo --workload c2 --format coo --steps 10 --no-cpu
run c5_coo --workload c5 --format coo --steps 10 --no-cpu
run c3_coo --workload c3 --format coo --steps 10 --no-cpu
run c3_css4 --workload c3 --format css --n-block 4 --steps 10 --no-cpu
run c3_jds --workload c3 --format jds --steps 10 --no-cpu
