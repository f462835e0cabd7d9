#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for b in 8 1; do
echo "=== B=$b"
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp10.py $b 2>&1 | grep -E "makespan|CTA 0|item [0-4]:|grid" | head -6 | cut -c1-700
done > gpurun_out/r2h_exp10.log 2>&1
cat gpurun_out/r2h_exp10.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "generator" 2>&1 | tail -3
timeout 300 python scratch/host_floor.py 2>&1 | grep -E "^(1|2|4|8|16|64) \{|dur" | cut -c1-330
