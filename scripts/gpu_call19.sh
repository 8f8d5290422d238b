source scripts/gpu_check.sh c19 > /dev/null 2>&1
run c2_css3 --workload c2 --steps 10 --no-cpu --no-also
B200SPMV_PERSIST=1 run c2_css3_persist --workload c2 --steps 10 --no-cpu --no-also
B200SPMV_PERSIST=1 run c2_css2_persist --workload c2 --format css --n-block 2 --steps 10 --no-cpu --no-also
B200SPMV_PERSIST=1 run c2_css4_persist --workload c2 --format css --n-block 4 --steps 10 --no-cpu --no-also
timeout 300 python bench.py --workload c5 --format coo --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_coo2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coo_tile_kernel -s 3 -c 1 -f -o gpurun_out/prof_coo_c5_v2 python bench.py --workload c5 --format coo --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_coo2.log 2>&1
echo "ncu coo rc=$?"
