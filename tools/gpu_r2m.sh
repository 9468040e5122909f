#!/bin/bash
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest.log
python - <<'PY'
import sys, os, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, bench
print(json.dumps({k: (round(v["ms"], 3), round(v["frac_of_sustained_bf16"], 3)) for k, v in bench.encoder_block(torch.device("cuda:0")).items()}))
PY
MDG_LINEAR_RES_DIRECT=1 python - <<'PY'
import sys, os, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, bench
print("direct residual loads:", json.dumps({k: (round(v["ms"], 3), round(v["frac_of_sustained_bf16"], 3)) for k, v in bench.encoder_block(torch.device("cuda:0")).items()}))
PY
