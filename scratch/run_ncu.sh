set -x
CMD="python bench.py --batch 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_loss_tc|k_gather_tc|k_dense_flat' -s 9 -c 3 -f -o gpurun_out/prof_r1_v3 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log
tail -5 gpurun_out/ncu_full.log
