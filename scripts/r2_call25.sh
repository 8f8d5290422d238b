# Round 2, GPU call 25 (2 GPUs): single-process path with the small-CTA pull kernel; torchrun path sanity after today's changes.
mkdir -p gpurun_out
timeout 300 python tests/mg_check.py 2 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_mg" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_check.py > gpurun_out/r2_dist_check_2.log 2>&1; echo "dist_check rc=$?"; grep -c "bit-identical=True host-step+fresh-x=True" gpurun_out/r2_dist_check_2.log
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']; print('N=2 GF %.1f ms %.4f exposed %.4f exchange %s timeouts %s launch %s' % (d['value'], d['ms_per_step'], c['exposed_exchange_ms'], c['exchange'], c.get('exchange_flag_timeouts'), c['launch'][:12]))"
