"""Per-rank check of the compressible gradient pool under torchrun: is the memory compressed on every device, and what
does each rank's step take with and without it?"""
import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps, timed_steps
rank, local = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
src, tgt = make_maps(LAYER_SETS['b5'], 64, torch.float32, dev, 1234 + rank)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    crit(src, tgt).backward()
res = {}
for comp in (True, False, True):
    pn.set_gradient_compression(comp)
    for _ in range(5): step()
    flag = all(pn.gradient_is_compressed(t.grad) for t in tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50): step()
    e1.record(); torch.cuda.synchronize()
    print(f'rank {rank} dev {local} compression={comp} compressed={flag} step {e0.elapsed_time(e1) / 50 * 1e3:.1f} us', flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
