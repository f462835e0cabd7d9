#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_netf_gpu.py tests/test_nhwc_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "head or netf or nhwc or channels or patch_sample" > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2u_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --head > gpurun_out/r2u_bench_head.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line --layout nhwc > gpurun_out/r2u_bench_nhwc.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2u_bench_head.json','gpurun_out/r2u_bench_nhwc.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k in ('ms_per_step','kernels_us'):
        print(k, json.dumps(d.get(k))[:700])
PY
