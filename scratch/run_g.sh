set -x
CMDH="python bench.py --head --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMDH > gpurun_out/plain_head.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_p\|k_wgrad_tc\|k_loss_tc_p -s 18 -c 6 -f -o gpurun_out/prof_r1_v6_head $CMDH > gpurun_out/ncu_full_head.log 2>&1
tail -3 gpurun_out/ncu_full_head.log
