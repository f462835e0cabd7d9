#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
timeout 300 python scratch/host_floor.py > gpurun_out/r2c_host_floor.log 2>&1; grep -v "^{" gpurun_out/r2c_host_floor.log | tail -30
