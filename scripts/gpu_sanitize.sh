# compute-sanitizer memcheck over the small parity cases (one tool per gpurun call, B200_PROFILING.md).
# Usage (next round): gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or empty_row or tile_boundaries or rejects" \
    > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_pytest.log; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/memcheck.log
