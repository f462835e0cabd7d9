#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','roofline_path','head_mode','strong','nccl_selfcheck','grad_reducer'):
    print(k, json.dumps(d.get(k))[:700])
n=d['nhwc']; print('nhwc', n['ms_per_step'], n['value'], n.get('strong'), n.get('head_mode',{}).get('ms_per_step'))
print('e2e', d['e2e']['value'], d['e2e'].get('channels_last'))
PY
