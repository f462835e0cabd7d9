"""Head-mode timing at bench size (B5, B=64, nc=256) and per-kernel breakdown."""
import os, sys, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda')
B = int(os.environ.get('B', '64'))
layers = LAYER_SETS['b5']
src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
torch.manual_seed(0)
netF = pn.PatchSampleF(use_mlp=True, nc=256).cuda()
netF.create_mlp(tgt)
math = os.environ.get('MATH', 'tc_bf16x3')
def step():
    for t in tgt: t.grad = None
    netF.zero_grad(set_to_none=True)
    loss, _ = pn.patchnce_with_head(netF, src, tgt, 0.07, 256, math=math)
    loss.backward()
    return loss
for _ in range(3): l = step()
torch.cuda.synchronize()
print('loss', l.item(), 'warnings', pn.poll_nonfinite_warnings(block=True))
n = 20
t0 = time.perf_counter()
for _ in range(n): step()
torch.cuda.synchronize()
print(f'head mode {math} B={B}: {(time.perf_counter()-t0)/n*1e3:.3f} ms/step')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
rows = [(e.key[:80], e.device_time_total / 5, e.count // 5) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1])[:14]:
    print(f'{t:10.1f} us/step x{c:3d}  {k}')
