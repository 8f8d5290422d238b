source scripts/gpu_check.sh c11
run c5_coo --workload c5 --format coo --steps 10 --no-cpu
run c2_coo --workload c2 --format coo --steps 10 --no-cpu
run c2_css3 --workload c2 --format css --n-block 3 --steps 10 --no-cpu
run c2_ell --workload c2 --format ell --steps 10 --no-cpu
run c4_dia --workload c4 --format dia --steps 10 --no-cpu
run c5_crs --workload c5 --format crs --steps 10 --no-cpu
