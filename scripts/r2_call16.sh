# Round 2, GPU call 16 (1 GPU): COO tests with both feeds, probe of one L2-resident column block.
mkdir -p gpurun_out
TAG=r2c16
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coo or guard_bands or hyb" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -2 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/r2_probe_block.py 2>&1 | tee gpurun_out/r2_probe_block.txt
for w in c3 c5 c4; do timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 --workload $w --format coo 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w coo auto: GF %.1f ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"; done
