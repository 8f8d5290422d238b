mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2c23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c23.log
tail -4 gpurun_out/pytest_r2c23.log; grep -E "^FAILED" gpurun_out/pytest_r2c23.log | head
