"""Per-kernel timings via torch.profiler (CUPTI) + a few what-if experiments."""
import os, sys, ctypes, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
lib = _lib.load()
dev = torch.device('cuda')
B = int(os.environ.get('B', '64'))
layers = LAYER_SETS['b5']
src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, math=os.environ.get('MATH', 'tc_bf16x3'))
gran = os.environ.get('L2GRAN')
if gran:
    lib.pnce_debug_set_l2_fetch_granularity.restype = ctypes.c_int
    print('L2 fetch granularity now', lib.pnce_debug_set_l2_fetch_granularity(int(gran)))

def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print(f'B={B} eager step {(time.perf_counter()-t0)/20*1e3:.3f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
rows = [(e.key[:60], e.device_time_total / e.count, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, n in sorted(rows, key=lambda r: -r[1])[:8]:
    print(f'{t:10.1f} us x{n:3d}  {k}')
cpu = [(e.key[:60], e.self_cpu_time_total / 10) for e in prof.key_averages()]
print('CPU self time per step (us), top:')
for k, t in sorted(cpu, key=lambda r: -r[1])[:14]:
    print(f'{t:10.1f}  {k}')
if os.environ.get('FILL'):
    n = sum(c*h*w for c,h,w,_ in layers) * B
    x = torch.empty(n, dtype=torch.float32, device=dev)
    for name, fn in (('zero_', lambda: x.zero_()), ('fill_(1)', lambda: x.fill_(1.0))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'{name}: {n*4/1e6:.0f} MB in {ms*1e3:.1f} us = {n*4/ms/1e6:.0f} GB/s')
