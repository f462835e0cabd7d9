#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_netf_gpu.py tests/test_abi_cpu.py -x -q > gpurun_out/r3h_tests.log 2>&1; tail -5 gpurun_out/r3h_tests.log
timeout 300 python scratch/split_breakdown.py 64 2>&1 | grep use_mlp | cut -c1-700
timeout 300 python scratch/split_breakdown.py 16 2>&1 | grep use_mlp | cut -c1-700
