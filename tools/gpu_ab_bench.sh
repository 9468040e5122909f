#!/bin/bash
# A/B of the working-tree library against libmadrigal_b200_prev.so (HEAD) on the whole bench line (step, encoder block)
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu 2>&1 | tail -1
for v in new prev new prev; do
  if [ $v = prev ]; then export MDG_LIB_PATH=$PWD/madrigal_b200/lib/libmadrigal_b200_prev.so; else unset MDG_LIB_PATH; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
e=d['encoder']
print('$v', 'step %.4f kern %.4f other %.4f | enc ' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['whole_step']['other_ms']) + ' '.join('%s %.2f ms (%.3f)' % (k.split('_')[-1], v['ms'], v['frac_of_sustained_bf16']) for k,v in e.items()), '| c2 %.3f' % d['config2_4096_x_963_single_gpu']['ms_per_step'])"
  python tools/time_topk.py 2>/dev/null | tail -3
done
