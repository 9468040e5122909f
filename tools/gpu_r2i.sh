#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_decoder_gpu.py tests/test_normalize_gpu.py -x -q -m gpu 2>&1 | tail -2
for kind in lut pwl; do KIND=$kind timeout 120 python tools/time_sym.py 2>&1 | tail -1; done
L=481 timeout 120 python tools/time_sym.py 2>&1 | tail -1
PACKED=1 timeout 120 python tools/time_sym.py 2>&1 | tail -1
