"""Host-side cost of one step at B=1 (the reference's default batch): cProfile of forward + backward."""
import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for _ in range(20): step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'B={B}: host {1e6*(t1-t0)/n:.1f} us/step issue time, {1e6*(t2-t0)/n:.1f} us/step incl. drain')
tf = tb = 0.0
for _ in range(n):
    for t in tgt: t.grad = None
    a = time.perf_counter(); loss = crit(src, tgt); b = time.perf_counter(); loss.backward(); c = time.perf_counter()
    tf += b - a; tb += c - b
print(f'forward call {1e6*tf/n:.1f} us, backward call {1e6*tb/n:.1f} us')
pr = cProfile.Profile(); pr.enable()
for _ in range(n): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
