#!/bin/bash
# Round-2 session B: pipelined vs legacy normaliser-layout epilogue, packed tiles.
mkdir -p gpurun_out
python -m pytest tests/test_decoder_gpu.py tests/test_normalize_gpu.py -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
for epi in legacy pipelined; do for kind in lut pwl; do
MDG_MIRROR_EPI=$epi KIND=$kind timeout 120 python tools/time_sym.py 2>&1 | tail -1
done; done
PACKED=1 KIND=lut timeout 120 python tools/time_sym.py 2>&1 | tail -1
PACKED=1 KIND=pwl timeout 120 python tools/time_sym.py 2>&1 | tail -1
for epi in legacy pipelined; do
MDG_MIRROR_EPI=$epi N=16384 L=32 timeout 120 python tools/time_sym.py 2>&1 | tail -1
MDG_MIRROR_EPI=$epi D=128 timeout 120 python tools/time_sym.py 2>&1 | tail -1
done
