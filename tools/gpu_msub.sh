#!/bin/bash
# A/B: 256-row tasks (default) vs 128-row tasks (MDG_MSUB=1) for the bench's rank kernel.
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block"
for m in 2 1 2 1; do
  MDG_MSUB=$m $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('msub=$m', 'step %.4f kern %.4f frac %.3f parity %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity_checked']))"
done
