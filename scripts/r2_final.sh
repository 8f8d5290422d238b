# Round 2 final single-GPU regression: the driver's own commands (pytest -m gpu, smoke, bench.py, bench.py --impl reference).
mkdir -p gpurun_out
TAG=r2final
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | grep real
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | grep real
python scripts/bench_table.py gpurun_out/bench_${TAG}_default.json gpurun_out/bench_${TAG}_reference.json > gpurun_out/r2_all_formats.md 2>&1; tail -60 gpurun_out/r2_all_formats.md
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=5 singlespmv_b200/plugin/bin/spmv_b200_crs synth:lap2d5:1024 > gpurun_out/driver_${TAG}_c1_crs_host.txt 2>&1; grep -E "Performance|KernelTime" gpurun_out/driver_${TAG}_c1_crs_host.txt
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=5 singlespmv_b200/plugin/bin/spmv_b200_crs_dev synth:lap3d7:512 > gpurun_out/driver_${TAG}_c5_crs_dev.txt 2>&1; grep -E "Performance|KernelTime|Roofline" gpurun_out/driver_${TAG}_c5_crs_dev.txt
