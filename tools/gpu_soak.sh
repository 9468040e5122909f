#!/bin/bash
# Soak: the GPU suite twice back to back and two more fuzz seeds (compute-sanitizer is closed on this pool; bad
# accesses are looked for with the guard-band tests instead).
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
for i in 1 2; do
  python -m pytest tests -q -m gpu > gpurun_out/soak_pytest_$i.log 2>&1; echo "pytest[$i] exit=$?"; tail -1 gpurun_out/soak_pytest_$i.log
done
for s in 7 8; do
  FUZZ_SEED=$s FUZZ_CASES=250 python tools/fuzz_decoder.py > gpurun_out/soak_fuzz_$s.log 2>&1; echo "fuzz[$s] exit=$?"; tail -1 gpurun_out/soak_fuzz_$s.log
done
