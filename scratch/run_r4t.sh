#!/bin/bash
# dense-backward kernels without pdl_enter (k_fill_zero had lost 8 % to the two instructions)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4t_tests.log 2>&1; tail -n 1 gpurun_out/r4t_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_20_bench_n1.json 2> gpurun_out/r4t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_20_bench_n1.json') if l.startswith('{')][-1])
print(d['steps'], d['ms_per_step'], d['roofline_path'], d['kernels_us'])
print('head', d['head_mode']['ms_per_step'], 'nhwc', d['nhwc']['ms_per_step'], d['nhwc']['roofline_path']['frac'], d['nhwc']['kernels_us'])
print('nhwc head', d['nhwc']['head_mode']['ms_per_step'], d['nhwc']['head_mode']['kernels_us'])
for c in d['configs']: print(c['config'], c['ms_per_step'], c.get('direct_ms_per_step'))
print(d['strong']['direct']['ms_per_step'])
PY
timeout 200 python scratch/pdl_ab.py 2>&1 | tail -n 1
