#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python scratch/cfg4_breakdown.py 2>&1 | grep -v Warn | tee gpurun_out/r2x_cfg4.log
