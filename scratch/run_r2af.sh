#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp40.py 2>&1 | grep -v -i warn | head -1 | tee gpurun_out/r2af_exp40.log
PNCE_EXPERIMENTS=1 timeout 600 python -m pytest tests/test_netf_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "head or netf or patch_sample" 2>&1 | tail -2
