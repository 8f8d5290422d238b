# Round 2, second ncu pass (one GPU): the kernels that changed after scripts/r2_ncu.sh -- COO entry stream (TMA-fed on c5, load-fed
# on c3), the sliced-ELL column blocks of c2.  Each ncu run follows a plain run of the same command line that exited 0.
mkdir -p gpurun_out
cap() { # tag, kernel regex, skip, count, bench args...
  tag=$1; k=$2; s=$3; c=$4; shift 4
  python bench.py "$@" > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/r2_prof_$tag python bench.py "$@" > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag rc=$?"; ls -la gpurun_out/r2_prof_$tag.ncu-rep 2>/dev/null | awk '{print $5}'
}
A="--steps 3 --warmup 3 --no-cpu"
python bench.py --workload c2 --format ell $A > gpurun_out/plain_launches_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c2_ell.csv python bench.py --workload c2 --format ell $A > gpurun_out/ncu_launches_c2.log 2>&1; echo "launch list c2 rc=$?"
python bench.py --workload c5 --format coo $A > gpurun_out/plain_launches_c5_coo.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c5_coo.csv python bench.py --workload c5 --format coo $A > gpurun_out/ncu_launches_c5_coo.log 2>&1; echo "launch list c5 coo rc=$?"
cap c5_coo coo_stream_kernel 3 1 --workload c5 --format coo $A
cap c3_coo coo_stream_kernel 3 1 --workload c3 --format coo $A
cap c2_ell ell_spmv_kernel 9 3 --workload c2 --format ell $A
cap c2_jds ell_spmv_kernel 9 3 --workload c2 --format jds $A
cap c2_ss ell_spmv_kernel 9 3 --workload c2 --format ss $A
cap c3_csr5 c5_compute_kernel 3 1 --workload c3 --format csr5 $A
