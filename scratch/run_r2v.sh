#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2v_tests.log
timeout 600 python scratch/stress.py 0 300 > gpurun_out/r2v_stress.log 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/r2v_stress.log
timeout 600 python scratch/stress2.py 0 80 > gpurun_out/r2v_stress2.log 2>&1; echo "stress2 rc=$?"; tail -4 gpurun_out/r2v_stress2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line > gpurun_out/r2v_bench.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --head > gpurun_out/r2v_bench_head.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line --layout nhwc > gpurun_out/r2v_bench_nhwc.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2v_bench.json','gpurun_out/r2v_bench_head.json','gpurun_out/r2v_bench_nhwc.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline_path']['frac'], json.dumps(d.get('kernels_us'))[:400])
PY
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp10.py 64 > gpurun_out/r2v_stamps.log 2>&1; head -8 gpurun_out/r2v_stamps.log | cut -c1-700
