source scripts/gpu_check.sh c3
for g in 32 64 128; do B200SPMV_L2_FETCH=$g run c2_ell_f$g --workload c2 --steps 10 --no-cpu; done
B200SPMV_L2_FETCH=32 run c5_crs_f32 --workload c5 --steps 10 --no-cpu
B200SPMV_L2_FETCH=32 run c3_crs_f32 --workload c3 --steps 10 --no-cpu
run c2_jds --workload c2 --format jds --steps 10 --no-cpu
B200SPMV_L2_FETCH=32 run c2_jds_f32 --workload c2 --format jds --steps 10 --no-cpu
run c4_dia --workload c4 --steps 20 --no-cpu
run c4_crs --workload c4 --format crs --steps 20 --no-cpu
run c4_ell --workload c4 --format ell --steps 20 --no-cpu
run c1_dia --workload c1 --format dia --steps 50 --no-cpu
run c1_ell --workload c1 --format ell --steps 50 --no-cpu
run c5_dia --workload c5 --format dia --steps 10 --no-cpu
timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_dia.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dia_spmv -s 3 -c 1 -o gpurun_out/prof_dia_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_dia.log 2>&1
timeout 300 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_crs.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_stream_kernel -s 3 -c 1 -o gpurun_out/prof_crs_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_crs.log 2>&1
