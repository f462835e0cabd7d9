"""Kernel timeline of one bench step (start / duration / gap) from the torch profiler."""
import sys, json, os, tempfile
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda'); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward()
for _ in range(10): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
f = os.path.join(tempfile.gettempdir(), 'trace.json'); prof.export_chrome_trace(f)
ev = [e for e in json.load(open(f))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])
# take the third step: find the 3rd k_prep
starts = [i for i, e in enumerate(ev) if 'k_prep' in e['name']]
a, b = starts[2] - 5, starts[3] - 5
t0 = ev[a]['ts']; prev_end = None
for e in ev[a:b]:
    gap = (e['ts'] - prev_end) if prev_end is not None else 0.0
    print(f"{e['ts']-t0:9.1f} us  dur {e['dur']:8.1f}  gap {gap:6.1f}  {e['name'][:70]}")
    prev_end = e['ts'] + e['dur']
print('step span', ev[b]['ts'] - t0)
