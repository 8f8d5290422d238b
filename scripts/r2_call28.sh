timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_mg or x_window or partitioned" 2>&1 | tail -1
timeout 100 python tests/mg_check.py 1 2>&1 | tail -1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
