# Multi-GPU check + bench.  Usage: bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tests/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep dist_check gpurun_out/dist_check_$N.log; tail -5 gpurun_out/dist_check_$N.log | grep -v dist_check | tail -3
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err; echo "bench rc=$?"; cat gpurun_out/bench_multi_$N.json; tail -3 gpurun_out/bench_multi_$N.err


