#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_multicopy_gpu.py -m gpu -x -q 2>&1 | tail -n 3
timeout 200 python scratch/graph_first_capture.py 1 > gpurun_out/r4g_graph1.log 2>&1; tail -n 30 gpurun_out/r4g_graph1.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r4g_bench.json 2> gpurun_out/r4g_bench.err; tail -c 300 gpurun_out/r4g_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4g_bench.json') if l.startswith('{')][-1])
print(json.dumps(d.get('train_step'), indent=1))
PY
