#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
timeout 400 python scratch/host_floor.py > gpurun_out/r2j_host_floor.log 2>&1; grep -v "^{" gpurun_out/r2j_host_floor.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline_path','kernels_us','head_mode','strong','configs','e2e','module_split'):
    print(k, json.dumps(d.get(k))[:900])
PY
