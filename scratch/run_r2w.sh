#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "more_than_256 or 1024 or p1024 or largest or half_precision_maps_head or odd_shapes or nhwc or channels" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2w_tests.log
timeout 600 python scratch/stress2.py 0 80 > gpurun_out/r2w_stress2.log 2>&1; echo "stress2 rc=$?"; tail -3 gpurun_out/r2w_stress2.log
timeout 300 python scratch/config_sweep.py > gpurun_out/r2w_sweep.log 2>&1; tail -15 gpurun_out/r2w_sweep.log
