source scripts/gpu_check.sh c13 > /dev/null 2>&1
for d in 0 1200 2400 4800; do
B200SPMV_PF=$d run c5_crs_pf$d --workload c5 --steps 10 --no-cpu
B200SPMV_PF=$d run c3_crs_pf$d --workload c3 --steps 10 --no-cpu
B200SPMV_PF=$d run c1_crs_pf$d --workload c1 --steps 50 --no-cpu
B200SPMV_PF=$d run c2_css3_pf$d --workload c2 --steps 10 --no-cpu --no-also
done
bash scripts/gpu_ncu_all.sh
