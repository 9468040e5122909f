#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests/test_decoder_gpu.py tests/test_normalize_gpu.py -x -q -m gpu 2>&1 | tail -1
for i in 1 2; do timeout 120 python tools/time_sym.py 2>&1 | tail -1; done
KIND=pwl timeout 120 python tools/time_sym.py 2>&1 | tail -1
MDG_MIRROR_EPI=legacy timeout 120 python tools/time_sym.py 2>&1 | tail -1
python tools/time_topk.py 2>&1 | tail -3
