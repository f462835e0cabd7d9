#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 0 31 23 27 29 30 15 16 3 0 31; do
  echo -n "PNCE_PDL=$v  "
  PNCE_PDL=$v timeout 200 python scratch/pdl_ab.py 2>&1 | tail -1
done > gpurun_out/r4c_pdl_ab.log 2>&1
cat gpurun_out/r4c_pdl_ab.log
