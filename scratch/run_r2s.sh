#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2s_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --head > gpurun_out/r2s_bench_head.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_head.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline_path','kernels_us'):
    print(k, json.dumps(d.get(k))[:700])
PY
