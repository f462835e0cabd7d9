"""Kernel breakdown of BASELINE config 4 (B=8, 512^2 maps, P=1024)."""
import sys
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda'); B = 8
shapes = [(64, 512, 512), (256, 128, 128), (256, 128, 128), (128, 256, 256), (64, 512, 512)]
g = torch.Generator(device=dev).manual_seed(1)
src = [torch.randn(B, *s, device=dev, generator=g) for s in shapes]
tgt = [torch.randn(B, *s, device=dev, generator=g).requires_grad_() for s in shapes]
for P in (1024, 512, 256):
    crit = pn.PatchNCELoss(0.07, P)
    def step():
        for t in tgt: t.grad = None
        loss = crit(src, tgt); loss.backward(); return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): step()
    e1.record(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {e.key: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    ks = '  '.join(f'{k.split("::")[-1][:14]}={v:.0f}' for k, v in rows.items() if 'pnce::k_' in k)
    print(f'P={P}: step {e0.elapsed_time(e1) / 50 * 1e3:.0f} us  [{ks}]', flush=True)
