source scripts/gpu_check.sh c9
for v in 3 4; do
B200SPMV_TS=$v run c5_crs_ts$v --workload c5 --steps 10 --no-cpu
B200SPMV_TS=$v run c1_crs_ts$v --workload c1 --steps 50 --no-cpu
B200SPMV_TS=$v run c3_crs_ts$v --workload c3 --steps 10 --no-cpu
B200SPMV_TS=$v run c4_crs_ts$v --workload c4 --format crs --steps 10 --no-cpu
B200SPMV_TS=$v run c2_css3_ts$v --workload c2 --format css --n-block 3 --steps 10 --no-cpu
done
run c3_cusparse --workload c3 --format csr5 --steps 10 --no-cpu --compare-cusparse
grep -o '"cusparse".*' gpurun_out/bench_c9_c3_cusparse.json | cut -c1-400
run c5_cusparse --workload c5 --format crs --steps 10 --no-cpu --compare-cusparse
grep -o '"cusparse".*' gpurun_out/bench_c9_c5_cusparse.json | cut -c1-400
