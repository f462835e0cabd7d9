"""cfg 4 (512^2 maps, P=1024, B=8): per-kernel breakdown."""
import sys
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
shapes = [(64, 512, 512), (256, 128, 128), (256, 128, 128), (128, 256, 256), (64, 512, 512)]
g = torch.Generator(device=dev).manual_seed(1)
src = [torch.randn(8, *s, device=dev, generator=g) for s in shapes]
tgt = [torch.randn(8, *s, device=dev, generator=g).requires_grad_() for s in shapes]
crit = pn.PatchNCELoss(0.07, 1024)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): step()
e1.record(); torch.cuda.synchronize()
print(f'cfg4 step {e0.elapsed_time(e1)/30*1e3:.1f} us')
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:7]:
    print(f'  {e.device_time_total/5:8.1f} us  {e.key[:70]}')
