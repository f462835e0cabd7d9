"""A/B of an L2 policy knob: python exp15.py <knob> (9: dxT evict_last, 10: gather streaming loads)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
KEY = int(sys.argv[1]) if len(sys.argv) > 1 else 9
for rep in range(2):
  for knob in (0, 1):
    lib.pnce_debug_set(KEY, knob)
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    n = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {e.key[:40]: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    print(f'knob{KEY}={knob}: step {ms*1e3:.1f} us; ' + '  '.join(f'{k.split("::")[-1][:12]}={v:.1f}' for k, v in rows.items() if 'pnce::k_' in k), f'loss {l.item():.6f}')
lib.pnce_debug_set(KEY, 1 if KEY == 9 else 0)
