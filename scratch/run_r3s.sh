#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl or dp or compress" > gpurun_out/r3s_tests.log 2>&1; tail -3 gpurun_out/r3s_tests.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3s_bench_n2.json 2> gpurun_out/r3s_bench_n2.err
tail -c 300 gpurun_out/r3s_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3s_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline_path']['frac'], d['config']['grad_memory'][:20], d.get('plain_alloc',{}).get('ms_per_step'))
print('head', d['head_mode']['ms_per_step'], d['head_mode'].get('collective'), 'strong', json.dumps(d['strong'])[:400])
print('nhwc', d['nhwc']['ms_per_step'], 'selfcheck', d.get('nccl_selfcheck'))
print('grad_reducer', d.get('grad_reducer'))
PY
