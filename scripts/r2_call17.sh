# Round 2, GPU call 17 (1 GPU): full GPU suite after the band heuristic, probe again, c2/c3 crs quick lines.
mkdir -p gpurun_out
TAG=r2c17
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/r2_probe_block.py 2>&1 | grep "crs" | tee gpurun_out/r2_probe_block2.txt
