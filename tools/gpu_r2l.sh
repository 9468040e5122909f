#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-encoder-block"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:pair_score_kernel<\(int\)3" -s 14 -c 1 -f -o gpurun_out/prof_gemm1 $B > gpurun_out/ncu_gemm1.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_gemm1.log
