#!/bin/bash
# Round-2 session A: tests, smoke, bench (PDL on / off), write-only bandwidth probe.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"
tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
MDG_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block > gpurun_out/bench_nopdl.log 2>&1; echo "bench nopdl exit=$?"
python - <<'PY'
import json
for f in ("bench.log", "bench_nopdl.log"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f kern %.4f frac %.3f whole %.3f parity %s e2e_ms %.2f" % (
            d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
            d["roofline"]["whole_step"]["frac"], d["parity_checked"], d["e2e"]["ms_per_step"]), d["clocks"])
        if "encoder" in d:
            for k, v in d["encoder"].items():
                print("  enc", k, "%.3f ms %.3g drugs/s %.0f TF frac %.3f launches %d" % (v["ms"], v["drugs_per_s"], v["tflops"], v["frac_of_sustained_bf16"], v["launches"]))
        if "cpu_baseline" in d:
            print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
./tools/probe/write_peak > gpurun_out/write_peak.txt 2>&1; cat gpurun_out/write_peak.txt
nproc; free -g | head -2; df -h /dev/shm | tail -1
