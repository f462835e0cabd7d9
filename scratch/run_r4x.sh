#!/bin/bash
# dry first pass in k_loss_tc_p (instruction-cache warm-up under the wait for the first logits tile)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4x_tests.log 2>&1; tail -n 1 gpurun_out/r4x_tests.log
timeout 200 python scratch/pdl_ab.py 2>&1 | tail -n 1
timeout 200 python scratch/pdl_ab.py 2>&1 | tail -n 1
timeout 300 python scratch/stress.py > gpurun_out/r4x_stress.log 2>&1; tail -n 1 gpurun_out/r4x_stress.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r4x_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4x_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['step_us']['p50'], d['kernels_us'], d['head_mode']['ms_per_step'], d['nhwc']['ms_per_step'], d['nhwc']['head_mode']['ms_per_step'])
for c in d['configs']: print(c['config'], c['ms_per_step'], c.get('direct_ms_per_step'))
PY
