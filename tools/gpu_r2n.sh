#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests/test_normalize_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/time_ensemble_fused.py 2>&1 | tail -7
