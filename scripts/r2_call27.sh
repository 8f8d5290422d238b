# Round 2, GPU call 27 (2 GPUs): single-process path chooses graph replay or eager launches at plan time.
timeout 200 python tests/mg_check.py 2 2>&1 | tail -1
B200SPMV_MG_GRAPH=1 timeout 200 python tests/mg_check.py 2 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_mg" 2>&1 | tail -1
