# Round 2, GPU call 6 (2 GPUs): full parity suite, single-process multi-GPU check (graph + eager), plugin mg binaries.
mkdir -p gpurun_out
TAG=r2c6
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
timeout 300 python tests/mg_check.py 2 2>&1 | tail -2
B200SPMV_MG_NO_GRAPH=1 timeout 300 python tests/mg_check.py 2 2>&1 | tail -1
timeout 300 python tests/mg_check.py 1 2>&1 | tail -1
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 B200_NGPU=2 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg_dev synth:lap3d7:256 > gpurun_out/driver_${TAG}_mg2_dev.txt 2>&1; grep -E "Performance|KernelTime|nGPU|Halo|Graph|invalid" gpurun_out/driver_${TAG}_mg2_dev.txt
SPMV_MIN_SECONDS=0.3 SPMV_NTRY=3 B200_NGPU=2 timeout 300 singlespmv_b200/plugin/bin/spmv_b200_crs_mg synth:lap3d7:256 > gpurun_out/driver_${TAG}_mg2_host.txt 2>&1; grep -E "Performance|KernelTime|nGPU|invalid|failed" gpurun_out/driver_${TAG}_mg2_host.txt
