#!/bin/bash
# A/B of the working-tree library against libmadrigal_b200_prev.so (built from HEAD): rank kernel, N = 4096 and 20000
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu -k "decoder or normalize" 2>&1 | tail -1
for v in new prev new prev; do
  if [ $v = prev ]; then export MDG_LIB_PATH=$PWD/madrigal_b200/lib/libmadrigal_b200_prev.so; else unset MDG_LIB_PATH; fi
  echo "== $v"
  SWEEP_N=${AB_N:-4096,20000} SWEEP_KINDS=${AB_KINDS:-lut,pwl,packed} python tools/rank_n_sweep.py 2>&1 | grep -v Warning | cut -c1-112
done
