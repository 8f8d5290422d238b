# Round 2, GPU call 18 (1 GPU): probe of a padded sliced-ELL per column block on config 2.
mkdir -p gpurun_out
timeout 900 python scripts/r2_probe_ellblocks.py 2>&1 | tail -5 | tee gpurun_out/r2_probe_ellblocks.txt
