source scripts/gpu_check.sh c6
run c3_csr5 --workload c3 --format csr5 --steps 10 --no-cpu
run c3_csr5_s32 --workload c3 --format csr5 --sigma 32 --steps 10 --no-cpu
run c3_csr5_s8 --workload c3 --format csr5 --sigma 8 --steps 10 --no-cpu
run c5_csr5 --workload c5 --format csr5 --steps 10 --no-cpu
run c2_csr5 --workload c2 --format csr5 --steps 10 --no-cpu
run c2_css2 --workload c2 --format css --n-block 2 --steps 10 --no-cpu
run c2_css3 --workload c2 --format css --n-block 3 --steps 10 --no-cpu
