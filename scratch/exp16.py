"""Why does bench.py's timed loop read ~20 us/step above the plain loop?  Bisect: NVML sampler thread, events."""
import sys
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps, ClockSampler
dev = torch.device('cuda'); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13], math=pn.DEFAULT_MATH)
torch.manual_seed(7)
evs = {i: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for i in range(0, 200, 4)}
def step(i=None, events=False):
    for t in tgt: t.grad = None
    loss = crit(src, tgt)
    if events and i in evs: evs[i][0].record()
    loss.backward()
    if events and i in evs: evs[i][1].record()
    return loss
def timed(n, sampler_period, events):
    for _ in range(5): step()
    torch.cuda.synchronize()
    s = ClockSampler(0, period=sampler_period) if sampler_period else None
    if s: s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n): step(i, events)
    e1.record(); torch.cuda.synchronize()
    if s: s.stop()
    return e0.elapsed_time(e1) / n * 1e3
for rep in range(2):
    for name, per, ev in (('plain', 0, False), ('events/4', 0, True), ('sampler 20ms', 0.02, False), ('sampler 5ms', 0.005, False), ('both', 0.02, True)):
        print(f'{name:14s} {timed(200, per, ev):8.1f} us/step')
