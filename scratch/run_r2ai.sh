#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for seed in 11 12; do timeout 900 python scratch/stress.py $seed 300 > gpurun_out/r2ai_stress_$seed.log 2>&1; echo "stress $seed rc=$?"; tail -4 gpurun_out/r2ai_stress_$seed.log; done
