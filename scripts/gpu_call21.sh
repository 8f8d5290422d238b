source scripts/gpu_check.sh c21
run c2_css3 --workload c2 --steps 20 --no-cpu --no-also
run c4_crs --workload c4 --format crs --steps 10 --no-cpu
run c5_coo --workload c5 --format coo --steps 10 --no-cpu
run c4_dia --workload c4 --steps 10 --no-cpu
