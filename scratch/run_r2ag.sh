#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ag_tests.log
timeout 600 python scratch/stress.py 1 200 > gpurun_out/r2ag_stress.log 2>&1; echo "stress rc=$?"; tail -2 gpurun_out/r2ag_stress.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line > gpurun_out/r2ag_bench.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --head > gpurun_out/r2ag_bench_head.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2ag_bench.json','gpurun_out/r2ag_bench_head.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline_path']['frac'], json.dumps(d.get('kernels_us'))[:400])
PY
timeout 300 python scratch/cfg4_breakdown.py 2>&1 | grep -v -i warn | head -2
