#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
for t in 8 16 24 48; do
MDG_TASKS_PER_CTA=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t$t.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/bench_t$t.log').read().strip().splitlines()[-1]);print('tasks/cta=$t', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
done
