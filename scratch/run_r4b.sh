#!/bin/bash
# programmatic dependent launch (launch_k / pdl_enter): parity, then A/B of the small-batch steps with PNCE_PDL=0 / 1
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4b_tests.log 2>&1; tail -2 gpurun_out/r4b_tests.log
for v in 0 1 0 1; do
  echo "== PNCE_PDL=$v"
  PNCE_PDL=$v timeout 300 python scratch/host_floor.py 2>&1 | tail -8
done > gpurun_out/r4b_host_floor.log 2>&1
for v in 0 1; do
  echo "== PNCE_PDL=$v"
  PNCE_PDL=$v timeout 300 python scratch/host_head.py 2>&1 | tail -12
done > gpurun_out/r4b_host_head.log 2>&1
for v in 0 1; do
  PNCE_PDL=$v python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r4b_bench_pdl$v.json 2>/dev/null
done
timeout 300 python scratch/stress.py > gpurun_out/r4b_stress.log 2>&1; tail -n 2 gpurun_out/r4b_stress.log
