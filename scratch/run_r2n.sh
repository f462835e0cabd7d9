#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_netf_gpu.py -m gpu -q > gpurun_out/r2n_netf_tests.log 2>&1; echo "netf tests rc=$?"; tail -30 gpurun_out/r2n_netf_tests.log
