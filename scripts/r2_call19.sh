# Round 2, GPU call 19 (1 GPU): sliced-ELL column-block engine -- tests, c2 formats.
mkdir -p gpurun_out
TAG=r2c19
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "column_block or css or golden or host_multiply or guard_bands" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
for f in css; do timeout 300 python bench.py --steps 20 --warmup 5 --workload c2 --format $f --n-block 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 $f: GF %.1f ms %.4f frac %.3f parity %s e2e %.2f ms convert %.0f ms' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'), d['e2e']['ms_per_step'], d['config']['convert_ms']))"; done

