"""Where do slow bench runs lose their time?  Per-step GPU durations (events) and host issue times."""
import sys, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda', 0); B = 64
torch.cuda.set_device(0)
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13], math=pn.DEFAULT_MATH)
torch.manual_seed(7)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for rep in range(3):
    for _ in range(5): step()
    torch.cuda.synchronize()
    n = 200
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    host = []
    ev[0].record()
    for i in range(n):
        t0 = time.perf_counter(); step(); host.append(time.perf_counter() - t0)
        ev[i + 1].record()
    torch.cuda.synchronize()
    d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n)]
    ds = sorted(d); hs = sorted(h * 1e6 for h in host)
    print(f'rep {rep}: mean {sum(d)/n:.1f} us; gpu/step p10 {ds[n//10]:.0f} p50 {ds[n//2]:.0f} p90 {ds[9*n//10]:.0f} max {ds[-1]:.0f}; '
          f'host p50 {hs[n//2]:.0f} p90 {hs[9*n//10]:.0f} max {hs[-1]:.0f}; first 8 steps {[round(x) for x in d[:8]]}')
