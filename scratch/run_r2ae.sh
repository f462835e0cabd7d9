#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --head"
timeout 300 $CMD > gpurun_out/r2ae_plain_head.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc_p' -s 13 -c 1 -f -o gpurun_out/r2_08_gemm_y $CMD > gpurun_out/r2ae_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2ae_ncu.log
