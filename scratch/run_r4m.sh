#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4m_tests.log 2>&1; tail -n 1 gpurun_out/r4m_tests.log
PNCE_FOLD_PREP=0 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4m_tests_nofold.log 2>&1; tail -n 1 gpurun_out/r4m_tests_nofold.log
for v in 0 1; do
  echo -n "PNCE_FOLD_PREP=$v  "
  PNCE_FOLD_PREP=$v timeout 200 python scratch/pdl_ab.py 2>&1 | tail -1
done > gpurun_out/r4m_fold_ab.log 2>&1
cat gpurun_out/r4m_fold_ab.log
timeout 300 python scratch/stress.py > gpurun_out/r4m_stress.log 2>&1; tail -n 1 gpurun_out/r4m_stress.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r4m_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4m_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['nhwc']['ms_per_step'], d['head_mode']['ms_per_step'])
for c in d['configs']: print(c['config'], c['ms_per_step'], c.get('direct_ms_per_step'))
PY
