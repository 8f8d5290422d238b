mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 120 python bench.py --workload c1 --steps 50 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 200 python bench.py --workload c2 --format crs --steps 10 --no-cpu > gpurun_out/bench_c2crs.json 2> gpurun_out/bench_c2crs.err
timeout 200 python bench.py --workload c5 --steps 10 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 200 python bench.py --workload c3 --steps 10 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ell_spmv -s 3 -c 1 -o gpurun_out/prof_ell_c2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ell.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench_c2.json
