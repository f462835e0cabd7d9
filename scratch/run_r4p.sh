#!/bin/bash
# end-of-round hardening: other stress seeds, the test-suite with programmatic dependent launch off
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python scratch/stress.py 1 300 > gpurun_out/r4p_stress_seed1.log 2>&1; grep -c MISMATCH gpurun_out/r4p_stress_seed1.log; tail -n 1 gpurun_out/r4p_stress_seed1.log
timeout 400 python scratch/stress.py 2 300 > gpurun_out/r4p_stress_seed2.log 2>&1; grep -c MISMATCH gpurun_out/r4p_stress_seed2.log; tail -n 1 gpurun_out/r4p_stress_seed2.log
timeout 400 python scratch/stress2.py 1 > gpurun_out/r4p_stress2_seed1.log 2>&1; tail -n 1 gpurun_out/r4p_stress2_seed1.log | cut -c1-200
PNCE_PDL=0 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 1
