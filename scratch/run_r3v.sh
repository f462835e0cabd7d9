#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3v_tests.log 2>&1; tail -3 gpurun_out/r3v_tests.log | cut -c1-300
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp50.py 64 nhwc 2>&1 | grep -v Warn | grep compress
