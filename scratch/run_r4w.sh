#!/bin/bash
# the driver's command at the final commit of the round
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_22_bench_n1_final.json 2> gpurun_out/r4w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_22_bench_n1_final.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value']/1e6, d['roofline_path'], d['step_us'])
print(d['roofline'])
print('head', d['head_mode']['ms_per_step'], 'nhwc', d['nhwc']['ms_per_step'], d['nhwc']['roofline_path']['frac'], 'nhwc head', d['nhwc']['head_mode']['ms_per_step'], d['nhwc']['head_mode'].get('roofline_path_frac'))
print('strong', d['strong']['ms_per_step'], d['strong']['direct']['ms_per_step'], 'module_split', d['module_split']['ms_per_step'], d['module_split']['list_form']['ms_per_step'])
for c in d['configs']: print(c['config'], c['ms_per_step'], c.get('direct_ms_per_step'), round(c['roofline_path']['frac'],3))
print('e2e', d['e2e']['value']/1e6, d['e2e']['with_grads_d2h']['value']/1e6, d['e2e']['channels_last']['value']/1e6, d['e2e']['bulk_copy']['value']/1e6)
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print(d['train_step']); print(d['next_rows']); print(d['clocks'])
PY
