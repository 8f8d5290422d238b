source scripts/gpu_check.sh c14
run c5_coo --workload c5 --format coo --steps 10 --no-cpu
run c3_coo --workload c3 --format coo --steps 10 --no-cpu
run c2_coo --workload c2 --format coo --steps 10 --no-cpu
run c3_csr5 --workload c3 --format csr5 --steps 10 --no-cpu
run c5_csr5 --workload c5 --format csr5 --steps 10 --no-cpu
run c1_csr5 --workload c1 --format csr5 --steps 50 --no-cpu
