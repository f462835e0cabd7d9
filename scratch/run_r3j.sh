#!/bin/bash
cd "$GRAFT_REPO_ROOT"
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err ) 2>&1 | grep real
tail -c 400 gpurun_out/r3j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline_path']['frac'], d['roofline_path'].get('frac_p50'))
print('module_split', json.dumps(d['module_split'])[:600])
for c in d['configs']: print(json.dumps({k:v for k,v in c.items() if k!='roofline_path'})[:300], c.get('roofline_path',{}).get('frac'))
print('head', d['head_mode']['ms_per_step'], 'nhwc', d['nhwc']['ms_per_step'], d['nhwc']['head_mode']['ms_per_step'])
print('e2e', d['e2e']['value'])
PY
