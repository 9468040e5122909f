#!/bin/bash
# Round-2 session D (2 GPUs): peer all-gather check, bench --gpus 2 (configs[2] strong scaling), GPU tests.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_allgather_test.py > gpurun_out/peer2.log 2>&1; echo "peer exit=$?"; tail -2 gpurun_out/peer2.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2.log 2>gpurun_out/bench2.err; echo "bench2 exit=$?"
tail -1 gpurun_out/bench2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=2 value %.4g ms %.4f kern %.4f frac %.3f whole %.3f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['parity']))
print('single', d.get('single_gpu_same_workload')); print('eff', d.get('strong_scaling_efficiency_vs_single_gpu_same_box')); print('e2e', d['e2e']); print('exchange', d.get('exchange'))
"; tail -5 gpurun_out/bench2.err
