#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python bench.py --batch 8 --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-head-line > gpurun_out/r2aa_b8.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aa_b8.json').read().strip().splitlines()[-1])
print('B=8 main', d['ms_per_step'], d['step_us'], d['kernels_us'])
PY
timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import torch, gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps, timed_steps, kernel_breakdown
dev=torch.device('cuda',0)
src,tgt=make_maps(LAYER_SETS['b5'],8,torch.float32,dev,1)
tgt=[t.requires_grad_() for t in tgt]
crit=pn.PatchNCELoss(0.07,256)
def step():
    for t in tgt: t.grad=None
    crit(src,tgt).backward()
print('before profiler', timed_steps(step,200,20,1,dev))
kernel_breakdown(step)
print('after profiler ', timed_steps(step,200,20,1,dev))
import torch.distributed as dist, os
os.environ.setdefault('MASTER_ADDR','127.0.0.1'); os.environ.setdefault('MASTER_PORT','29577')
dist.init_process_group('nccl', rank=0, world_size=1, device_id=dev)
x=torch.ones(4,device=dev); dist.all_reduce(x)
print('after nccl init', timed_steps(step,200,20,1,dev))
PY
