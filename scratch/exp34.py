"""A/B: item order of the persistent loss kernel (debug knob 11: -1 = heavy first, 0 = one wave of light items first) x
gather order (knob 8)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
ref = None
for rep in range(2):
  for rot, gord in ((-1, 0), (24, 0), (48, 0), (64, 0), (74, 0), (84, 0), (100, 0), (124, 0), (74, 1), (202, 0), (222, 0)):
    lib.pnce_debug_set(11, rot); lib.pnce_debug_set(8, gord)
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    n = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    torch.manual_seed(7); l = step(); sig = (l.item(), [float(t.grad.double().abs().sum()) for t in tgt])
    if ref is None: ref = sig
    assert sig == ref, ('results changed', sig, ref)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {e.key: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    ks = '  '.join(f'{k.split("::")[-1][:12]}={v:.1f}' for k, v in rows.items() if 'pnce::k_' in k)
    print(f'loss order {rot:4d} gather-in-layer-order {gord}: step {ms*1e3:.1f} us  [{ks}]', flush=True)
lib.pnce_debug_set(11, 0); lib.pnce_debug_set(8, 0)
