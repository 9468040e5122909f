#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -k "pipelined or packed or symmetric" > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
for kind in lut pwl; do
MDG_MIRROR_EPI=pipelined KIND=$kind timeout 120 python tools/time_sym.py 2>&1 | tail -1
done
for t in 4 8 12 24 32; do
echo "tasks_per_cta=$t"; MDG_TASKS_PER_CTA=$t MDG_MIRROR_EPI=legacy timeout 120 python tools/time_sym.py 2>&1 | tail -1
MDG_TASKS_PER_CTA=$t MDG_MIRROR_EPI=pipelined timeout 120 python tools/time_sym.py 2>&1 | tail -1
done
