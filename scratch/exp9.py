"""A/B: persistent vs one-CTA-per-item loss kernel (debug knob 6), kernel times from the torch profiler."""
import sys, ctypes, time
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
for knob, ctas in ((1, 0), (0, 0), (0, 296 // 2), (0, 132)):
    lib.pnce_debug_set(6, knob); lib.pnce_debug_set(7, ctas)
    torch.manual_seed(3)
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    torch.manual_seed(3); l = step(); g = tgt[1].grad.clone()
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
    lk = [v for k, v in rows.items() if 'k_loss_tc' in k]
    print(f'no_persist={knob} ctas={ctas or "SMs"}: step {ms*1e3:.1f} us, loss kernel {lk[0]:.1f} us, loss {l.item():.6f}, grad sum {g.double().abs().sum().item():.9e}, warnings {pn.poll_nonfinite_warnings(block=True)}')
lib.pnce_debug_set(6, 0); lib.pnce_debug_set(7, 0)
