#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp40.py 2>&1 | grep -v -i warn | tee gpurun_out/r2ac_exp40.log
