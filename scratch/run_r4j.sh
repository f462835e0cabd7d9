#!/bin/bash
# capture 17: the headline command at the end-of-round code (programmatic dependent launch on), launch list + ncu --set full
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/r4j_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_17_final_launches.csv $CMD > gpurun_out/r4j_ncu_launches.log 2>&1
$CMD > gpurun_out/r4j_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p\|k_gather_tc\|k_dense_flat\|k_prep -s 12 -c 4 -f -o gpurun_out/r2_17_final $CMD > gpurun_out/r4j_ncu_full.log 2>&1
tail -2 gpurun_out/r4j_ncu_full.log | cut -c1-200
python scratch/ncu_summary.py gpurun_out/r2_17_final.ncu-rep > gpurun_out/r2_17_final_summary.md 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_17_bench_n1.json 2> gpurun_out/r4j_bench.err
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
