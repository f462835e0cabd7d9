#!/bin/bash
# is the seed-1 mismatch (case 269, tau = 0.015) tied to programmatic dependent launch / the folded prep, or a numerical edge?
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "0 0" "23 1"; do
  set -- $cfg
  echo "== PNCE_PDL=$1 PNCE_FOLD_PREP=$2"
  STRESS_ID_SEED=1 STRESS_VERBOSE=1 PNCE_PDL=$1 PNCE_FOLD_PREP=$2 timeout 400 python scratch/stress.py 1 300 2>&1 | grep -v Warn | tail -n 12
done > gpurun_out/r4q_stress_ab.log 2>&1
cat gpurun_out/r4q_stress_ab.log
