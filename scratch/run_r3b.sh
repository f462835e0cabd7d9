#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -m gpu -x -q -k "head or netf or nccl or dp" > gpurun_out/r3b_tests.log 2>&1; tail -4 gpurun_out/r3b_tests.log
timeout 300 python scratch/host_head.py 1 8 16 64 > gpurun_out/r3b_host_head.log 2>&1
grep -E "^B=|run_backward|patchnce.py" gpurun_out/r3b_host_head.log | cut -c1-200
