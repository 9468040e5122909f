#!/bin/bash
# ncu evidence for one bench step: launch list (shares) + full capture of the dominant kernel.  One GPU, after a plain run.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-encoder-block"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k "regex:pair_score_kernel|fused_encoder_kernel|convert_z_kernel|convert_w_kernel|convert_rows_kernel|peer_allgather" -c 2000 \
    --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
if [ "$1" == "full" ]; then
$B > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:pair_score_kernel<\(int\)6" -s 4 -c 1 -f -o gpurun_out/prof_rank $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_full.log
fi
