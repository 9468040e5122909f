#!/bin/bash
# Round 2, session 2, call 1: GPU suite at HEAD (incl. the chemCPA mlp dosers), rank-kernel sweep over the catalogue size,
# selected ncu metrics of one N = 20,000 launch.
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
SECONDS=0
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$? (${SECONDS}s)"; tail -3 gpurun_out/pytest.log
SECONDS=0
timeout 600 python tools/rank_n_sweep.py > gpurun_out/n_sweep.txt 2>&1; echo "sweep exit=$? (${SECONDS}s)"; cat gpurun_out/n_sweep.txt
SECONDS=0
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__cycles_active.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,smsp__issue_active.avg.per_cycle_active
for n in 4096 20000; do
  l=$([ $n = 4096 ] && echo 86 || echo 4)
  BIG_N=$n BIG_L=$l timeout 600 ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:pair_score_kernelILi6ELi8E --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/ncu_rank_n$n.csv python tools/big_once.py > gpurun_out/ncu_rank_n$n.log 2>&1
  echo "ncu n=$n exit=$? (${SECONDS}s)"; grep -v "^==" gpurun_out/ncu_rank_n$n.csv | cut -d, -f5,13- | head -20
done
