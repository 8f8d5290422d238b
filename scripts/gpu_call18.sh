source scripts/gpu_check.sh c18 > /dev/null 2>&1
for f in crs coo ell jds dia ss css csr5; do run c5_$f --workload c5 --format $f --steps 10 --no-cpu; done
for f in crs coo ell jds dia ss csr5; do run c4_$f --workload c4 --format $f --steps 10 --no-cpu; done
for f in crs coo ell jds ss csr5; do run c3_$f --workload c3 --format $f --steps 10 --no-cpu; done
timeout 300 python -m pytest tests/test_sweep.py -m gpu -x -q 2>&1 | tail -3
