#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3k_tests.log 2>&1; tail -15 gpurun_out/r3k_tests.log | cut -c1-300
for c in 1 0; do
PNCE_GRAD_COMPRESSION=$c timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r3k_bench_c$c.json 2> gpurun_out/r3k_bench_c$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3k_bench_c$c.json').read().strip().splitlines()[-1])
print('compression=$c', d['value'], d['ms_per_step'], d['roofline_path']['frac'], d['kernels_us'])
print('  nhwc', d['nhwc']['ms_per_step'], d['nhwc']['kernels_us'], 'head', d['head_mode']['ms_per_step'], 'nhwc head', d['nhwc']['head_mode']['ms_per_step'])
for c in d['configs']: print('  ', c['config'], c.get('ms_per_step'), c.get('direct_ms_per_step'), c.get('error'))
PY
done
