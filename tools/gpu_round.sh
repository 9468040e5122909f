#!/bin/bash
# One GPU-box session: tests, smoke, bench, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"
tail -1 gpurun_out/bench.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
$B > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_score_kernel -s 28 -c 2 -f -o gpurun_out/prof_pair $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
tail -3 gpurun_out/pytest.log
