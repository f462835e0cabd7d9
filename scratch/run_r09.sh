set -x
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_v9.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p\|k_gather_tc\|k_dense_flat\|k_prep -s 12 -c 4 -f -o gpurun_out/prof_r1_v9 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
