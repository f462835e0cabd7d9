#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2fin_tests_n2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2fin_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2fin_bench_n2.json 2> gpurun_out/r2fin_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2fin_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline_path']['frac'], 'head', d['head_mode']['ms_per_step'], 'strong', d['strong']['ms_per_step'], 'nhwc', d['nhwc']['ms_per_step'], d['nhwc']['head_mode']['ms_per_step'], d['nccl_selfcheck']['ok'])
PY
