#!/bin/bash
# 2-GPU box: the reducer through NCCL with the multi-tensor pack / unpack, bench N=2 (grad_reducer line), train_step line on one GPU
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py tests/test_dp_multicopy_gpu.py -m gpu -q > gpurun_out/r4i_tests.log 2>&1; tail -n 2 gpurun_out/r4i_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r4i_bench_n2.json 2> gpurun_out/r4i_bench_n2.err
tail -c 300 gpurun_out/r4i_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4i_bench_n2.json') if l.startswith('{')][-1])
print(json.dumps(d.get('grad_reducer')))
PY
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r4i_bench_n1.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4i_bench_n1.json') if l.startswith('{')][-1])
print(json.dumps(d.get('train_step'), indent=1))
PY
