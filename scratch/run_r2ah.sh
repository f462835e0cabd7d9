#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_netf_gpu.py tests/test_parity_gpu.py tests/test_dp_nccl_gpu.py -m gpu -q -x -k "head or netf or patch_sample" > gpurun_out/r2ah_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ah_tests.log
timeout 600 python scratch/stress2.py 3 60 > gpurun_out/r2ah_stress2.log 2>&1; echo "stress2 rc=$?"; tail -2 gpurun_out/r2ah_stress2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --head > gpurun_out/r2ah_bench_head.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2ah_bench_head.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline_path']['frac'], json.dumps(d.get('kernels_us'))[:400])
PY
