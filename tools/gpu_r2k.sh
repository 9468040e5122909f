#!/bin/bash
mkdir -p gpurun_out
ls -la madrigal_b200/lib/
python -c "
from madrigal_b200 import build
print('stamp', open(build.STAMP_PATH).read().strip()[:16], 'now', build.source_hash()[:16])
"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.4g ms %.4f kern %.4f frac %.3f whole %.3f other %.4f parity %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
    d["roofline"]["whole_step"]["frac"], d["roofline"]["whole_step"]["other_ms"], d["parity_checked"]))
PY
python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -1
