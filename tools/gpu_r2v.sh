#!/bin/bash
# Decomposition of the rank kernel at N = 4096 and N = 20000: full / no look-ups (store path) / no stores (look-ups + MMA)
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
for d in 0 1 2 0 1 2; do
  echo "== MDG_DEBUG_EPI=$d (0 full, 1 no look-ups, 2 no stores)"
  MDG_DEBUG_EPI=$d SWEEP_N=4096,20000 SWEEP_KINDS=lut,packed python tools/rank_n_sweep.py 2>&1 | grep -v Warning
done
