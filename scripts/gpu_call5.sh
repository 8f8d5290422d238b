source scripts/gpu_check.sh c5 > /dev/null 2>&1
for m in 0 1 2 3 4 5; do B200SPMV_XLOAD=$m run c2_ell_x$m --workload c2 --steps 10 --no-cpu; done
M="lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_lookup_miss.sum"
for m in 0 1 3; do
B200SPMV_XLOAD=$m timeout 600 ncu --metrics $M --clock-control none -k regex:ell_spmv -s 3 -c 1 --csv --log-file gpurun_out/ncu_ell_x$m.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ell_x$m.log 2>&1
done
