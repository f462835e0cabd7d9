#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3n_tests.log 2>&1; tail -3 gpurun_out/r3n_tests.log | cut -c1-300
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; tail -c 300 gpurun_out/r3n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline_path'], d['roofline']['frac'], d['roofline']['launch_ms'])
print(d['config']['grad_memory'][:40], d.get('plain_alloc'))
print('nhwc', d['nhwc']['ms_per_step'], d['nhwc']['roofline_path']['frac'], 'head', d['head_mode']['ms_per_step'], d['nhwc']['head_mode']['ms_per_step'], 'strong', d['strong']['ms_per_step'])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line"
$CMD > gpurun_out/r3n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3_01_launches.csv $CMD > gpurun_out/r3n_ncu_launches.log 2>&1
$CMD > gpurun_out/r3n_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loss_tc_p\|k_gather_tc\|k_dense_flat\|k_prep -s 12 -c 4 -f -o gpurun_out/r3_01_compressed $CMD > gpurun_out/r3n_ncu_full.log 2>&1
tail -2 gpurun_out/r3n_ncu_full.log | cut -c1-200
