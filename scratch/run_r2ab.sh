#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ab_bench.json').read().strip().splitlines()[-1])
print(sorted(d.keys()))
for k in ('value','ms_per_step','roofline','roofline_path','kernels_us','clocks','gpu_launches','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:500])
print('head', d['head_mode']['ms_per_step'], 'strong', d['strong']['ms_per_step'])
n=d['nhwc']; print('nhwc', n['ms_per_step'], n['roofline_path']['frac'], n.get('kernels_us'), n['head_mode'])
print('configs', json.dumps(d['configs'])[:1500])
print('module_split', d['module_split'])
print('e2e', d['e2e']['value'], d['e2e']['channels_last'], d['e2e']['with_grads_d2h']['value'], d['e2e']['bulk_copy']['value'])
PY
