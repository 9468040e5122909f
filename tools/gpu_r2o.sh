#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
for g in 1 2 4; do
echo "task group $g"; MDG_TASK_GROUP=$g timeout 120 python tools/time_sym.py 2>&1 | tail -1
MDG_TASK_GROUP=$g MDG_TASKS_PER_CTA=32 timeout 120 python tools/time_sym.py 2>&1 | tail -1
done
MDG_TASK_GROUP=2 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -1
