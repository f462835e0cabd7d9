#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_netf_gpu.py -m gpu -q -x > gpurun_out/r2m_netf_tests.log 2>&1; echo "netf tests rc=$?"; tail -40 gpurun_out/r2m_netf_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_netf_gpu.py > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2m_tests.log
