#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -5 gpurun_out/r2f_tests.log
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp10.py 64 2>&1 | grep -E "makespan|CTA 0|item [0-4]:" | head -7 | cut -c1-700 > gpurun_out/r2f_exp10.log; cat gpurun_out/r2f_exp10.log
timeout 600 python scratch/stress.py 0 200 > gpurun_out/r2f_stress.log 2>&1; tail -4 gpurun_out/r2f_stress.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline_path'], d['kernels_us'], d['step_us'])
PY
