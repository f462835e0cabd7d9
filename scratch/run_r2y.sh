#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python scratch/cfg4_host.py 2>&1 | grep -v -i warn | tee gpurun_out/r2y_cfg4_host.log
