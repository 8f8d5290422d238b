source scripts/gpu_check.sh c22 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in 1 2 4 8; do for sl in 0 1; do
B200SPMV_HOST_CHUNKS=$c B200SPMV_HOST_SLICES=$sl run c2_css3_ch${c}_sl$sl --workload c2 --steps 10 --no-cpu --no-also
done; done
B200SPMV_HOST_CHUNKS=1 run c2_ell_ch1 --workload c2 --format ell --steps 10 --no-cpu
B200SPMV_HOST_CHUNKS=8 run c2_ell_ch8 --workload c2 --format ell --steps 10 --no-cpu
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -4
