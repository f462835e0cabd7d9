#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 23 31 27; do echo "== PNCE_PDL=$v"; PNCE_PDL=$v timeout 200 python scratch/pdl_timeline.py 2>&1 | grep -v Warn | tail -28; done > gpurun_out/r4e_pdl_timeline.log 2>&1
cat gpurun_out/r4e_pdl_timeline.log
