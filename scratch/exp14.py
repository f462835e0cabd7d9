"""A/B: dense backward staged in shared memory (flags 0) vs store-first (flags 2); fill ceilings (flags 1, 3)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = 64
for dt in (torch.float32,):
    src, tgt = make_maps(LAYER_SETS['b5'], B, dt, dev, 1234)
    tgt = [t.requires_grad_() for t in tgt]
    crit = pn.PatchNCELoss(0.07, 256)
    def step():
        for t in tgt: t.grad = None
        loss = crit(src, tgt); loss.backward(); return loss
    ref = None
    for flags in (0, 8, 16, 1):
        lib.pnce_debug_set(1, flags)
        torch.manual_seed(3)
        for _ in range(5): l = step()
        torch.cuda.synchronize()
        torch.manual_seed(3); step(); g = [t.grad.clone() for t in tgt]
        if flags == 0 and ref is None: ref = g
        same = all(torch.equal(a, b) for a, b in zip(g, ref)) if flags in (0, 4) else None
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10): step()
            torch.cuda.synchronize()
        rows = {e.key[:40]: e.device_time_total / 10 for e in prof.key_averages() if e.device_time_total > 0 and 'k_dense' in e.key}
        print(f'{dt} flags={flags}: ' + '  '.join(f'{k.split("::")[-1][:16]}={v:.1f}us' for k, v in rows.items()), 'grads identical to staged:', same)
    lib.pnce_debug_set(1, 0)
    del src, tgt; torch.cuda.empty_cache()
