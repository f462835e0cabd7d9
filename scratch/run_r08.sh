set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py --head --no-e2e --no-cpu-baseline > gpurun_out/bench_head.json 2>/dev/null
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_v8.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p\|k_gather_tc\|k_dense_flat\|k_prep -s 12 -c 4 -f -o gpurun_out/prof_r1_v8 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
CMD2="python scratch/prof_next_rows.py"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_amp\|k_aug\|k_hinge\|k_multi --csv --log-file gpurun_out/launches_r1_v8_next.csv $CMD2 > gpurun_out/ncu_launches2.log 2>&1
$CMD2 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_amp\|k_aug\|k_hinge\|k_multi -s 11 -c 11 -f -o gpurun_out/prof_r1_v8_next $CMD2 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-200
cat gpurun_out/bench_default.json | cut -c1-300
