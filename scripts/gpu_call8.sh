source scripts/gpu_check.sh c8
for v in 1 2; do
B200SPMV_TS=$v run c5_crs_ts$v --workload c5 --steps 10 --no-cpu
B200SPMV_TS=$v run c1_crs_ts$v --workload c1 --steps 50 --no-cpu
B200SPMV_TS=$v run c3_crs_ts$v --workload c3 --steps 10 --no-cpu
B200SPMV_TS=$v run c4_crs_ts$v --workload c4 --format crs --steps 10 --no-cpu
B200SPMV_TS=$v run c2_css3_ts$v --workload c2 --format css --n-block 3 --steps 10 --no-cpu
done
run c4_dia --workload c4 --steps 20 --no-cpu
run c5_dia --workload c5 --format dia --steps 10 --no-cpu
run c1_dia --workload c1 --format dia --steps 50 --no-cpu
