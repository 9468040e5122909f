#!/bin/bash
# Store-path experiments of the rank kernel (mirror layout): which part of the 2.5-3.1 ms store path is DRAM, which is SM?
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
run() { echo "== MDG_DEBUG_EPI=$1 : $2"; MDG_DEBUG_EPI=$1 SWEEP_N=4096,20000 SWEEP_KINDS=lut python tools/rank_n_sweep.py 2>&1 | grep -v Warning; }
run 1 "no look-ups (store path)"
run 5 "no look-ups, all stores into a 2 MB L2-resident window"
run 9 "no look-ups, plain stores only (no transposed store)"
run 17 "no look-ups, transposed stores only"
run 33 "no look-ups, evict_first"
run 65 "no look-ups, evict_last"
run 128 "full kernel, evict_normal hint"
run 32 "full kernel, evict_first"
run 64 "full kernel, evict_last"
run 0 "full kernel"
