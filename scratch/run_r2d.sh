#!/bin/bash
# round 2, call D: single-pass softmax WIP -- tests, stress, bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
timeout 600 python scratch/stress.py 0 300 > gpurun_out/r2d_stress.log 2>&1; tail -8 gpurun_out/r2d_stress.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 2500 gpurun_out/r2d_bench.json
