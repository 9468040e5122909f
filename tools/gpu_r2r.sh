#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/time_gemm1.py; MDG_GEMM1_NARROW_STORE=1 python tools/time_gemm1.py
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g ms %.4f kern %.4f frac %.3f whole %.3f other %.4f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['roofline']['whole_step']['other_ms'], d['parity_checked']))"
