#!/bin/bash
# Round-2 final single-GPU check: GPU tests, smoke, default bench, reference arm.
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.4g ms %.4f kern %.4f frac %.3f whole %.3f parity %s e2e %.2f packed %.2f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["parity_checked"], d["e2e"]["ms_per_step"], d["e2e_packed_tiles"]["ms_per_step"], d["gpu_launches"]))
print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
PY
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-160
