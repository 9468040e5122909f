#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
echo skip-tests
python tools/time_ensemble_fused.py 2>&1 | tail -7
