# Round 2, GPU call 22 (1 GPU): full GPU suite with the CRS entry stream, c3 crs line with parity, ncu capture of it.
mkdir -p gpurun_out
TAG=r2c22
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload c3 --format crs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 crs: GF %.1f ms %.4f frac %.3f parity %s kernel %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'), d['roofline']['kernel']))"
A="--steps 3 --warmup 3 --no-cpu"
python bench.py --workload c3 --format crs $A > gpurun_out/plain_c3_crs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:entry_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_c3_crs python bench.py --workload c3 --format crs $A > gpurun_out/ncu_c3_crs.log 2>&1; echo "ncu rc=$?"
