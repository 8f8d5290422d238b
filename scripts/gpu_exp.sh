source scripts/gpu_check.sh exp > /dev/null 2>&1
for m in 0 1; do for it in 4 8 16; do
B200SPMV_RBS_MODE=$m B200SPMV_RBS_ITERS=$it run c5_crs_m${m}_i$it --workload c5 --format crs --steps 10 --no-cpu
done; done
B200SPMV_RBS_MODE=0 run c1_crs_m0 --workload c1 --format crs --steps 50 --no-cpu
B200SPMV_RBS_MODE=1 run c1_crs_m1 --workload c1 --format crs --steps 50 --no-cpu
