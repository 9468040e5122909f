#!/bin/bash
mkdir -p gpurun_out
MDG_BENCH_CONFIG3_DRYRUN=2304,12 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-encoder-block > gpurun_out/bench_dry.log 2>gpurun_out/bench_dry.err; echo "dry exit=$?"
tail -1 gpurun_out/bench_dry.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps(d.get('config3_20k_x_953'), indent=1))
"; tail -5 gpurun_out/bench_dry.err | cut -c1-300
