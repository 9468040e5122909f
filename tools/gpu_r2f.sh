#!/bin/bash
mkdir -p gpurun_out
for epi in legacy pipelined; do
MDG_MIRROR_EPI=$epi N=20000 L=119 timeout 300 python tools/time_sym.py 2>&1 | tail -1
done
python tools/time_pair_score.py 2>&1 | grep -E "logit|sigmoid"
python -m pytest tests/test_fusion_gpu.py tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
