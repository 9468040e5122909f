#!/bin/bash
# One GPU-box session: tests, smoke, bench (+ reference arm), ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"
tail -1 gpurun_out/bench.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
if [ "$1" == "ncu" ]; then
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
fi
