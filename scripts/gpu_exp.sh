source scripts/gpu_check.sh exp > /dev/null 2>&1
for u in 3 5 7 9; do
B200SPMV_DIA_U=$u run c4_dia_u$u --workload c4 --steps 20 --no-cpu
B200SPMV_DIA_U=$u run c5_dia_u$u --workload c5 --format dia --steps 10 --no-cpu
B200SPMV_DIA_U=$u run c1_dia_u$u --workload c1 --format dia --steps 50 --no-cpu
done
