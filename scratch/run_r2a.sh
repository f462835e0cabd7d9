#!/bin/bash
# round 2, call A: full GPU test suite + host floor + bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 300 python scratch/host_floor.py > gpurun_out/r2a_host_floor.log 2>&1; tail -8 gpurun_out/r2a_host_floor.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 3000 gpurun_out/r2a_bench.json
