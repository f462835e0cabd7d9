#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2aj_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2aj_tests.log
timeout 600 python scratch/stress.py 21 200 > gpurun_out/r2aj_stress.log 2>&1; echo "stress rc=$?"; tail -2 gpurun_out/r2aj_stress.log
timeout 400 python scratch/host_floor.py 2>&1 | grep -E "^(1|2|4|8|16|64) |k_prep" | head -12
timeout 300 python scratch/cfg4_breakdown.py 2>&1 | grep -v -i warn | head -1
