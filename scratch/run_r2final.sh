#!/bin/bash
# end-of-round capture: ncu --set full + launch list of the headline command on the round's last commit
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
timeout 300 $CMD > gpurun_out/r2f_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_loss_tc_p|k_gather_tc|k_dense_flat|k_prep' -s 12 -c 4 -f -o gpurun_out/r2_09_final $CMD > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2f_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 40 --csv --log-file gpurun_out/r2_09_launches.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
CMDH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --head --layout nhwc"
timeout 300 $CMDH > gpurun_out/r2f_plain_head.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_loss_tc_p|k_gemm_tc_p|k_wgrad_tc|k_gather_tc_nhwc|k_dense_nhwc' -s 24 -c 8 -f -o gpurun_out/r2_10_head_nhwc $CMDH > gpurun_out/r2f_ncu_head.log 2>&1
echo "ncu head rc=$?"; tail -2 gpurun_out/r2f_ncu_head.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 60 --csv --log-file gpurun_out/r2_10_head_nhwc_launches.csv $CMDH > /dev/null 2>&1; echo "launch list head rc=$?"
