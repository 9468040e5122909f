#!/bin/bash
# host-mirror e2e: parity test + bench e2e legs (plain / host mirror / packed)
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu -k "packed or host" > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
nproc; lscpu | grep -i "model name"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-encoder-block > gpurun_out/bench_mirror.log 2>gpurun_out/bench_mirror.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_mirror.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_mirror.log").read().strip().splitlines()[-1])
print("step %.4f kern %.4f frac %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
for k in ("e2e", "e2e_host_mirror", "e2e_packed_tiles"):
    print(k, {a: b for a, b in d[k].items() if a != "note"})
PY
python - <<'PY'
import torch, time
from madrigal_b200 import decoder
N=4096; L=10
T = decoder.packed_tiles_per_outcome(N)
packed = torch.randint(0, 30000, (L, T, 32, 32), dtype=torch.int16).view(torch.uint16).pin_memory()
out = torch.empty((L, N, N), dtype=torch.uint16).pin_memory()
decoder.mirror_packed_tiles_host(packed, N, out=out, threads=8)
for th in (1, 4, 8, 16, 32):
    t0 = time.perf_counter()
    for _ in range(3): decoder.mirror_packed_tiles_host(packed, N, out=out, threads=th)
    dt = (time.perf_counter()-t0)/3
    print(th, "threads: %.2f ms -> %.1f GB/s out" % (dt*1e3, out.numel()*2/dt/1e9))
PY
