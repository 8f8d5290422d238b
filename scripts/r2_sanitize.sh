# compute-sanitizer over the small parity cases, ONE tool per gpurun call (B200_PROFILING.md).
# Usage: gpurun --timeout 1500 -- 'bash scripts/r2_sanitize.sh memcheck|racecheck|synccheck'
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL="golden or short_row or column_blocked or hyb or css_blocks or dense_row or tile_boundaries or chunk_boundaries or empty_row or rejects or multiply_rows or fp32_variant or col_extent or x_window or row_blocked or test_coo"
if [ "$TOOL" = racecheck ]; then SEL="crs_golden or coo_tile_boundaries or chunk_boundaries or short_row or column_blocked_layout or css_blocks or test_hyb or dia_golden or csr5_empty or x_window"; fi
timeout 1300 compute-sanitizer --tool $TOOL --error-exitcode 3 --log-file gpurun_out/r2_$TOOL.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_${TOOL}_pytest.log 2>&1
echo "$TOOL rc=$?"; tail -3 gpurun_out/r2_${TOOL}_pytest.log; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/r2_$TOOL.log; grep "ERROR SUMMARY" gpurun_out/r2_$TOOL.log | sort | uniq -c | head; grep -m5 -A12 "Invalid\|Race\|hazard" gpurun_out/r2_$TOOL.log | head -60
