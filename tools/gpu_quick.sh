#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
python tools/time_topk.py 2>&1 | tail -4
python tools/time_pair_score.py 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
python tools/time_encoder.py 2>&1 | tail -8
