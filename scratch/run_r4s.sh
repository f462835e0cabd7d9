#!/bin/bash
# final state of the round: smoke, every GPU test, the default bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r4s_tests.log 2>&1; tail -n 1 gpurun_out/r4s_tests.log
python bench.py > gpurun_out/r2_19_bench_n1_default.json 2> gpurun_out/r4s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_19_bench_n1_default.json') if l.startswith('{')][-1])
print(d['steps'], d['ms_per_step'], d['roofline_path'], d['kernels_us'])
print(d['head_mode']['ms_per_step'], d['nhwc']['ms_per_step'], d['nhwc']['head_mode']['kernels_us'])
print(d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
PY
