# Round 2, GPU call 14 (1 GPU): gather + stream microbenchmark, guard-band tests, index64 leg, COO final.
mkdir -p gpurun_out
TAG=r2c14
python - <<'PY'
import ctypes as C, os
cmp = C.CDLL(os.path.join("singlespmv_b200", "libb200cmp.so"))
for tb in (45 << 20, 67 << 20, 134217728):
    ms = C.c_float()
    mg = cmp.b200cmp_gather(C.c_longlong(tb), C.c_longlong(1 << 29), 2, 5, C.byref(ms))
    ms2 = C.c_float()
    mg2 = cmp.b200cmp_gather_stream(C.c_longlong(tb), C.c_longlong(1 << 29), 2, 5, C.byref(ms2))
    print("table %4d MB: gathers alone %.3f ms per 536.9M (%.0f G/s); with 12 B/gather stream %.3f ms (%.0f G/s, stream %.0f GB/s)"
          % (tb >> 20, ms.value * 536.870912 / mg, mg * 1e6 / (ms.value * 1e-3) / 1e9, ms2.value * 536.870912 / mg2, mg2 * 1e6 / (ms2.value * 1e-3) / 1e9,
             12 * mg2 * 1e6 / (ms2.value * 1e-3) / 1e9))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "guard_bands or coo" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --mini --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_mini.json 2> gpurun_out/bench_${TAG}_mini.err; echo "mini rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2c14_mini.json"))
print("index64 (mini):", d.get("index64"))
PY
