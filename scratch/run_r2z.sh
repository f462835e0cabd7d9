#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py tests/test_nhwc_gpu.py -m gpu -q > gpurun_out/r2z_nccl_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_nccl_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2z_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','roofline_path','head_mode','strong','nhwc','nccl_selfcheck','grad_reducer'):
    print(k, json.dumps(d.get(k))[:900])
print('e2e', json.dumps(d['e2e'])[:300], json.dumps(d['e2e'].get('channels_last')))
PY
