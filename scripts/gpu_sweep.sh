# Every format on the stencil / skewed configs + the small and the headline config (one GPU, ~5 min).
source scripts/gpu_check.sh sweep
for f in crs coo ell jds dia ss css csr5; do run c5_$f --workload c5 --format $f --steps 10 --no-cpu; done
for f in crs coo ell jds dia ss csr5; do run c4_$f --workload c4 --format $f --steps 10 --no-cpu; done
for f in crs coo ss csr5; do run c3_$f --workload c3 --format $f --steps 10 --no-cpu; done
for f in crs dia ell jds csr5; do run c1_$f --workload c1 --format $f --steps 50 --no-cpu; done
run c2_css --workload c2 --steps 20 --no-cpu
