#!/bin/bash
# Round-2 session E: tests after the burst-load epilogues, bench, configs[3] single-GPU slice (legacy vs pipelined).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.4g ms %.4f kern %.4f frac %.3f whole %.3f parity %s e2e_ms %.2f e2e_eq %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
    d["roofline"]["whole_step"]["frac"], d["parity_checked"], d["e2e"]["ms_per_step"], d["parity"]["e2e_output_equals_device_output"]), d["clocks"])
for k, v in d.get("encoder", {}).items():
    print("  enc", k, "%.3f ms %.3g drugs/s %.0f TF frac %.3f launches %d" % (v["ms"], v["drugs_per_s"], v["tflops"], v["frac_of_sustained_bf16"], v["launches"]))
PY
tail -3 gpurun_out/bench.err
for epi in legacy pipelined; do
MDG_MIRROR_EPI=$epi N=20000 L=119 timeout 300 python tools/time_sym.py 2>&1 | tail -1
done
python tools/time_pair_score.py 2>&1 | tail -8
