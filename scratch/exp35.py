"""Split forward (debug knob 12): loss of the heavy layers on n CTAs of a second stream beside the gather of the light layers."""
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
def fwd():
    with torch.no_grad(): return crit(src, tgt)
ref = None
for rep in range(2):
  for split in (0, 32, 48, 64, 74, 90, 110):
    lib.pnce_debug_set(12, split)
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    n = 100
    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_step, t_fwd = timed(step), timed(fwd)
    torch.manual_seed(7); l = step(); sig = (l.item(), [float(t.grad.double().abs().sum()) for t in tgt])
    if ref is None: ref = sig
    assert sig == ref, ('results changed', sig, ref)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type.name == 'CUDA' and 'pnce::k_' in e.name:
            rows.setdefault(e.name.split('::')[-1][:12], []).append(e.device_time)
    ks = '  '.join(f'{k}={"/".join(str(round(sum(v[i::len(v)//5]) / 5)) for i in range(len(v)//5))}' for k, v in rows.items())
    print(f'split {split:3d}: step {t_step:.1f} us, forward alone {t_fwd:.1f} us  [{ks}]', flush=True)
lib.pnce_debug_set(12, 0)
