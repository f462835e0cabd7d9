#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_netf_gpu.py tests/test_nhwc_gpu.py -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2o_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','head_mode','nhwc','module_split'):
    print(k, json.dumps(d.get(k))[:1800])
PY
