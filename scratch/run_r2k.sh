#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nhwc_gpu.py -m gpu -q -x > gpurun_out/r2k_nhwc_tests.log 2>&1; echo "nhwc tests rc=$?"; tail -25 gpurun_out/r2k_nhwc_tests.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_nhwc_gpu.py > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline_path','kernels_us','nhwc'):
    print(k, json.dumps(d.get(k))[:1500])
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layout nhwc --no-head-line > gpurun_out/r2k_bench_nhwc.json 2> gpurun_out/r2k_bench_nhwc.err; echo "bench nhwc rc=$?"; tail -3 gpurun_out/r2k_bench_nhwc.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_nhwc.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','roofline_path','kernels_us','e2e'):
    print(k, json.dumps(d.get(k))[:1200])
PY
