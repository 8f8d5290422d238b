# Round 2 ncu pass (one GPU): launch lists + full captures of the dominant kernels.  Each ncu run follows a plain run of the
# same command line that exited 0 (B200_PROFILING.md).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
cap() { # tag, kernel regex, skip, count, bench args...
  tag=$1; k=$2; s=$3; c=$4; shift 4
  python bench.py "$@" > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/r2_prof_$tag python bench.py "$@" > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag rc=$?"; ls -la gpurun_out/r2_prof_$tag.ncu-rep 2>/dev/null | awk '{print $5}'
}
A="--steps 3 --warmup 3 --no-cpu"
python bench.py --workload c5 $A > gpurun_out/plain_launches_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c5_default.csv python bench.py --workload c5 $A > gpurun_out/ncu_launches_c5.log 2>&1; echo "launch list c5 rc=$?"
python bench.py --workload c2 --format ell $A > gpurun_out/plain_launches_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c2_ell.csv python bench.py --workload c2 --format ell $A > gpurun_out/ncu_launches_c2.log 2>&1; echo "launch list c2 rc=$?"
cap c5_crs chunk_stream_kernel 3 1 --workload c5 $A
cap c1_crs chunk_stream_kernel 3 1 --workload c1 $A
cap c2_ell tile_stream_kernel 9 3 --workload c2 --format ell $A
cap c2_jds tile_stream_kernel 9 3 --workload c2 --format jds $A
cap c3_crs tile_stream_kernel 3 1 --workload c3 --format crs $A
cap c4_crs tile_stream_kernel 3 1 --workload c4 --format crs $A
cap c5_coo coo_tile_kernel 3 1 --workload c5 --format coo $A
cap c5_crs_f32 chunk_stream_kernel 3 1 --workload c5 --precision 1 $A
