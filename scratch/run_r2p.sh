#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp10.py 64 > gpurun_out/r2p_stamps.log 2>&1; cat gpurun_out/r2p_stamps.log | cut -c1-900
