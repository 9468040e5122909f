#!/bin/bash
# Round-2 session G (8 GPUs): bench --gpus 8 = configs[2] strong scaling + configs[3] leg, peer all-gather check.
mkdir -p gpurun_out
echo skip-peer-test
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8.log 2>gpurun_out/bench8.err; echo "bench8 exit=$?"
tail -1 gpurun_out/bench8.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=8 value %.4g ms %.4f kern %.4f frac %.3f whole %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac']))
print('parity', d['parity']); print('single', d.get('single_gpu_same_workload')); print('eff', d.get('strong_scaling_efficiency_vs_single_gpu_same_box')); print('e2e', d['e2e']); print('exchange', d.get('exchange'))
print('config3', json.dumps(d.get('config3_20k_x_953'), indent=1))
"; tail -5 gpurun_out/bench8.err | cut -c1-400
