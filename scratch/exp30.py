"""Can a zero fill run UNDER the loss kernel?  k_loss_tc_p holds 226+1 KB of shared memory and 59 k registers per SM: one
128-thread CTA without shared memory still fits beside it.  A side stream waits for the post-gather event (debug knob 10)
and zero-fills a buffer while the loss kernel runs; compare with forward alone and fill alone."""
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
ev = torch.cuda.Event(); ev.record(); torch.cuda.synchronize()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def fwd():
    with torch.no_grad():
        return crit(src, [t.detach() for t in tgt])
for gb in (0.27, 0.54, 0.81, 1.07):
    Z = torch.empty(int(gb * 2**30) // 4, dtype=torch.float32, device=dev)
    def fill(): Z.zero_()
    def both():
        lib.pnce_debug_set(10, ev.cuda_event)
        l = fwd()
        lib.pnce_debug_set(10, 0)
        side.wait_event(ev)
        with torch.cuda.stream(side): Z.zero_()
        main.wait_stream(side)
        return l
    def serial():
        l = fwd(); Z.zero_(); return l
    t_f, t_z, t_s, t_b = timed(fwd), timed(fill), timed(serial), timed(both)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): both()
        torch.cuda.synchronize()
    rows = {e.key: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    ks = '  '.join(f'{k.split("::")[-1][:14]}={v:.0f}' for k, v in rows.items() if 'k_loss' in k or 'k_gather' in k or 'Fill' in k or 'fill' in k)
    print(f'fill {gb:.2f} GiB: fwd {t_f:.0f} us, fill {t_z:.0f} us, serial {t_s:.0f} us, overlapped {t_b:.0f} us (hidden {t_s - t_b:.0f} us)  [{ks}]', flush=True)
