#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.4g ms %.4f kern %.4f frac %.3f whole %.3f parity %s e2e %.2f packed %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["parity_checked"], d["e2e"]["ms_per_step"], d["e2e_packed_tiles"]["ms_per_step"]))
print("config2 single:", d.get("config2_4096_x_963_single_gpu"))
print("cpu:", d.get("cpu_baseline"))
print("keys:", sorted(d.keys()))
PY
tail -3 gpurun_out/bench.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
