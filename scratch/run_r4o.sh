#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4o_tests.log 2>&1; tail -n 1 gpurun_out/r4o_tests.log
for v in 0 1; do
  echo "== PNCE_FOLD_PREP=$v"
  PNCE_FOLD_PREP=$v timeout 300 python scratch/host_head.py 2>&1 | grep "head step\|head direct"
done > gpurun_out/r4o_host_head.log 2>&1
cat gpurun_out/r4o_host_head.log
timeout 300 python scratch/stress2.py > gpurun_out/r4o_stress2.log 2>&1; tail -n 1 gpurun_out/r4o_stress2.log | cut -c1-120
