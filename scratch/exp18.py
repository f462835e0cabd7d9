"""Module-split API (PatchSampleF -> PatchNCELoss(feat_q, feat_k)) vs the fused call, same maps."""
import sys, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
samp = pn.PatchSampleF(use_mlp=False)
crit = pn.PatchNCELoss(0.07, 256)
def step_split():
    for t in tgt: t.grad = None
    with torch.no_grad():
        fk, ids = samp(src, 256, None)
    fq, _ = samp(tgt, 256, ids)
    loss = sum(crit(q, k, batch_size=B) for q, k in zip(fq, fk)) / len(fq)
    loss.backward(); return loss
def step_fused():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for name, fn in (('fused', step_fused), ('module split', step_split)):
    torch.manual_seed(1)
    for _ in range(3): l = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1)/20:.3f} ms/step, loss {l.item():.5f}')
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    rows = [(e.key[:70], e.device_time_total / 3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
    for k, t, c in sorted(rows, key=lambda r: -r[1])[:8]:
        print(f'   {t:9.1f} us/step x{c:3d}  {k}')
