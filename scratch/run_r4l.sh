#!/bin/bash
# id prep folded into the gather for small problems (k_gather_tc_fold): parity, A/B, stress
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4l_tests.log 2>&1; tail -n 2 gpurun_out/r4l_tests.log
for v in 0 1 0 1; do
  echo -n "PNCE_FOLD_PREP=$v  "
  PNCE_FOLD_PREP=$v timeout 200 python scratch/pdl_ab.py 2>&1 | tail -1
done > gpurun_out/r4l_fold_ab.log 2>&1
cat gpurun_out/r4l_fold_ab.log
timeout 300 python scratch/stress.py > gpurun_out/r4l_stress.log 2>&1; tail -n 1 gpurun_out/r4l_stress.log
timeout 300 python scratch/stress2.py > gpurun_out/r4l_stress2.log 2>&1; tail -n 1 gpurun_out/r4l_stress2.log | cut -c1-150
