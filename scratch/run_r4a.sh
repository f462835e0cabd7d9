#!/bin/bash
# capture 14: channels-last path at the end-of-round code (two-launch backward: k_fill_zero + k_scatter_nhwc)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --layout nhwc --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/r4a_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_14_nhwc_launches.csv $CMD > gpurun_out/r4a_ncu_launches.log 2>&1
$CMD > gpurun_out/r4a_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p\|k_gather_tc_nhwc\|k_fill_zero\|k_scatter_nhwc\|k_prep -s 15 -c 5 -f -o gpurun_out/r2_14_nhwc $CMD > gpurun_out/r4a_ncu_full.log 2>&1
tail -2 gpurun_out/r4a_ncu_full.log | cut -c1-200
python scratch/ncu_summary.py gpurun_out/r2_14_nhwc.ncu-rep > gpurun_out/r2_14_nhwc_summary.md 2>&1
python bench.py --layout nhwc --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line > gpurun_out/r4a_bench_nhwc.json 2>/dev/null
# stress at the final code
timeout 400 python scratch/stress.py > gpurun_out/r4a_stress.log 2>&1; tail -3 gpurun_out/r4a_stress.log
timeout 400 python scratch/stress2.py > gpurun_out/r4a_stress2.log 2>&1; tail -3 gpurun_out/r4a_stress2.log
