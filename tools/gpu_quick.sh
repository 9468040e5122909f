#!/bin/bash
# Quick GPU check after a small change: stale-library guard, the GPU suite, smoke.
mkdir -p gpurun_out
python -c "from madrigal_b200 import build; import sys; sys.exit(0 if build.library_is_current() else 1)" || { echo "STALE LIBRARY"; exit 1; }
python -m pytest tests -x -q -m gpu "$@" > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
