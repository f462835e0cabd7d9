#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p -s 6 -c 1 -f -o gpurun_out/prof_r2_g $CMD > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log | cut -c1-200
