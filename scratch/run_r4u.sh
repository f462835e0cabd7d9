#!/bin/bash
# final code on a 2-GPU box: every GPU test (NCCL ones included)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_21_gpu_tests_2gpu_box_final.log 2>&1; tail -n 2 gpurun_out/r2_21_gpu_tests_2gpu_box_final.log
