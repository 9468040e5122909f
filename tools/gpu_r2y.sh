#!/bin/bash
# evict_first hint on the rank stores vs catalogue size (full kernel)
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
for d in 0 32 0 32; do
  echo "== MDG_DEBUG_EPI=$d (0 default, 32 evict_first)"
  MDG_DEBUG_EPI=$d SWEEP_N=4096,6144,8192,12288,16384,20000 SWEEP_KINDS=lut,pwl python tools/rank_n_sweep.py 2>&1 | grep -v Warning
done
