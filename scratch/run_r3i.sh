#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_netf_gpu.py tests/test_nhwc_gpu.py -x -q > gpurun_out/r3i_tests.log 2>&1; tail -3 gpurun_out/r3i_tests.log
timeout 300 python scratch/split_breakdown.py 64 2>&1 | grep use_mlp | cut -c1-200
timeout 300 python scratch/split_breakdown.py 16 prof > gpurun_out/r3i_prof.log 2>&1; grep use_mlp gpurun_out/r3i_prof.log | cut -c1-200
