#!/bin/bash
# Round-2 final multi-GPU validation: bench.py --gpus $1 through torchrun, plus the reference arm's quiet non-zero ranks.
N=${1:-4}
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench$N.log 2>gpurun_out/bench$N.err; echo "bench$N exit=$?"
tail -1 gpurun_out/bench$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=%d value %.4g ms %.4f kern %.4f frac %.3f whole %.3f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac']))
p = d['parity']; print('parity', p['checked'], p.get('checksum_equals_single_gpu'), p['e2e_output_equals_device_output'])
s = d.get('single_gpu_same_workload'); print('single', s and s['ms_per_step'], 'eff', d.get('strong_scaling_efficiency_vs_single_gpu_same_box'))
print('e2e', d['e2e']['ms_per_step'], d['e2e']['sample'], 'packed', d['e2e_packed_tiles']['ms_per_step'], d['e2e_packed_tiles']['unpacked_equals_device_output']); print('exchange', d.get('exchange', {}).get('mode'))
c = d.get('config3_20k_x_953')
if c: print('config3', c['rank_u16']['ms_per_step'], c['roofline']['frac'], c['roofline']['kernel_frac_rank0'], c['top1000']['ms_per_step'], c['parity_checked'])
"; grep -v "^W1018\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench$N.err | tail -5 | cut -c1-300
