"""Where does a bench step spend its time? CPU vs GPU, fwd vs bwd."""
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
crit = pn.PatchNCELoss(0.07, 256)
n = 50
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
cpu_f, cpu_b = [], []
def run(record):
    for i in range(n):
        for t in tgt: t.grad = None
        if record: ev[i][0].record()
        t0 = time.perf_counter()
        loss = crit(src, tgt)
        t1 = time.perf_counter()
        if record: ev[i][1].record()
        loss.backward()
        t2 = time.perf_counter()
        if record: ev[i][2].record()
        if record: cpu_f.append(t1 - t0); cpu_b.append(t2 - t1)
run(False); torch.cuda.synchronize()
t0 = time.perf_counter(); run(True); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n
med = lambda x: sorted(x)[len(x)//2]
print(f'wall/step {wall*1e3:.3f} ms; GPU fwd {med([e[0].elapsed_time(e[1]) for e in ev]):.3f} ms, GPU bwd {med([e[1].elapsed_time(e[2]) for e in ev]):.3f} ms; '
      f'CPU fwd call {med(cpu_f)*1e3:.3f} ms, CPU bwd call {med(cpu_b)*1e3:.3f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(False); torch.cuda.synchronize()
rows = [(e.key[:70], e.device_time_total / e.count, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1])[:12]:
    print(f'{t:10.1f} us x{c:4d}  {k}')
import ctypes
from gan_variant_research_b200 import _lib
lib = _lib.load()
lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
for chunks in (1, 2, 3):
    lib.pnce_debug_set(5, chunks)
    cpu_f.clear(); cpu_b.clear()
    run(False); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(True); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n
    print(f'fwd_chunks knob={chunks}: wall/step {wall*1e3:.3f} ms; GPU fwd {med([e[0].elapsed_time(e[1]) for e in ev]):.3f} ms, GPU bwd {med([e[1].elapsed_time(e[2]) for e in ev]):.3f} ms')
