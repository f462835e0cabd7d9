#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4d_tests.log 2>&1; tail -n 2 gpurun_out/r4d_tests.log
python bench.py > gpurun_out/r4d_bench_pdl23.json 2> gpurun_out/r4d_bench.err
PNCE_PDL=0 python bench.py --no-cpu-baseline > gpurun_out/r4d_bench_pdl0.json 2>> gpurun_out/r4d_bench.err
timeout 300 python scratch/stress2.py > gpurun_out/r4d_stress2.log 2>&1; tail -n 2 gpurun_out/r4d_stress2.log
