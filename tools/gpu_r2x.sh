#!/bin/bash
# full ncu capture (with source) of one launch of the pipelined rank kernel at configs[1] size
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
export BIG_N=4096 BIG_L=86
python tools/big_once.py > gpurun_out/plain_big.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pair_score_kernelILi6ELi8E -s 1 -c 1 -f -o gpurun_out/prof_rank2 python tools/big_once.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_full2.log; ls -la gpurun_out/*.ncu-rep
