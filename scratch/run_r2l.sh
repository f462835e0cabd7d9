#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nhwc_gpu.py -m gpu -q > gpurun_out/r2l_nhwc_tests.log 2>&1; echo "nhwc tests rc=$?"; tail -25 gpurun_out/r2l_nhwc_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --head > gpurun_out/r2l_bench_head.json 2> gpurun_out/r2l_bench_head.err; echo "bench head rc=$?"; tail -3 gpurun_out/r2l_bench_head.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_head.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline_path','kernels_us'):
    print(k, json.dumps(d.get(k))[:1500])
PY
