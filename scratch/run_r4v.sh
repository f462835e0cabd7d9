#!/bin/bash
# griddepcontrol.wait behind the tcgen05 kernels' prologue (barrier init, TMEM allocation)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4v_tests.log 2>&1; tail -n 1 gpurun_out/r4v_tests.log
timeout 200 python scratch/pdl_ab.py 2>&1 | tail -n 1
timeout 300 python scratch/host_head.py 2>&1 | grep "head step\|head direct"
timeout 300 python scratch/stress2.py > gpurun_out/r4v_stress2.log 2>&1; tail -n 1 gpurun_out/r4v_stress2.log | cut -c1-100
timeout 300 python scratch/stress.py > gpurun_out/r4v_stress.log 2>&1; tail -n 1 gpurun_out/r4v_stress.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r4v_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4v_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['head_mode']['ms_per_step'], d['nhwc']['ms_per_step'], d['nhwc']['head_mode']['ms_per_step'])
for c in d['configs']: print(c['config'], c['ms_per_step'], c.get('direct_ms_per_step'))
PY
