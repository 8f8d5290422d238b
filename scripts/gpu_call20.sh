source scripts/gpu_check.sh c20
run c4_crs --workload c4 --format crs --steps 10 --no-cpu
run c3_crs --workload c3 --format crs --steps 10 --no-cpu
run c5_crs --workload c5 --format crs --steps 10 --no-cpu
run c2_ss --workload c2 --format ss --steps 10 --no-cpu
run c2_css3 --workload c2 --steps 10 --no-cpu --no-also
run c5_coo --workload c5 --format coo --steps 10 --no-cpu
run c4_coo --workload c4 --format coo --steps 10 --no-cpu
run c3_coo --workload c3 --format coo --steps 10 --no-cpu
