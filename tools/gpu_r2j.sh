#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY: rebuild before gpurun"; exit 1; }
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.4g ms %.4f kern %.4f frac %.3f whole %.3f other %.4f parity %s e2e_ms %.2f packed_ms %.2f packed_ok %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
    d["roofline"]["whole_step"]["frac"], d["roofline"]["whole_step"]["other_ms"], d["parity_checked"], d["e2e"]["ms_per_step"],
    d["e2e_packed_tiles"]["ms_per_step"], d["e2e_packed_tiles"]["unpacked_equals_device_output"]), d["clocks"])
for k, v in d.get("encoder", {}).items():
    print("  enc", k, "%.3f ms %.3g drugs/s %.0f TF frac %.3f launches %d" % (v["ms"], v["drugs_per_s"], v["tflops"], v["frac_of_sustained_bf16"], v["launches"]))
PY
tail -3 gpurun_out/bench.err
python tools/time_topk.py 2>&1 | tail -3
