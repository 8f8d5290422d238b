source scripts/gpu_check.sh exp > /dev/null 2>&1
B200SPMV_RBS_MAXLEN=32 run c4_crs_rbs32 --workload c4 --format crs --steps 10 --no-cpu
B200SPMV_RBS_MAXLEN=32 run c2_ss_rbs32 --workload c2 --format ss --steps 10 --no-cpu
run c5_ss --workload c5 --format ss --steps 10 --no-cpu
