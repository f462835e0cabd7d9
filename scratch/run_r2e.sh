#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for k in 1 0 2; do
echo "=== knob9=$k"
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp10.py 64 0 $k 2>&1 | grep -E "makespan|CTA 0|item [12]:" | head -4 | cut -c1-900
done > gpurun_out/r2e_exp10.log 2>&1
cat gpurun_out/r2e_exp10.log
