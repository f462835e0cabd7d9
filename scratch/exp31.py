"""exp30 with control: the forward on a HIGH-priority stream (its loss CTAs are placed first), the fill on a normal one,
with a fill kernel of our own (128 threads, no shared memory, max shared carve-out) as persistent or flat grid."""
import sys, ctypes
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
lib.pnce_debug_fill.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
dev = torch.device('cuda'); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
crit = pn.PatchNCELoss(0.07, 256)
ev = torch.cuda.Event(); ev.record(); torch.cuda.synchronize()
hi = torch.cuda.Stream(priority=-1); lo = torch.cuda.Stream(priority=0)
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def fwd():
    with torch.no_grad(): return crit(src, tgt)
gb = 0.54
Z = torch.empty(int(gb * 2**30) // 4, dtype=torch.float32, device=dev)
main = torch.cuda.current_stream()
for ctas in (148, 296, 592, 0):
    def fill(stream=None):
        st = (stream or torch.cuda.current_stream()).cuda_stream
        assert lib.pnce_debug_fill(Z.data_ptr(), Z.numel() * 4, ctas, st) == 0
    def both():
        hi.wait_stream(main)
        with torch.cuda.stream(hi):
            lib.pnce_debug_set(10, ev.cuda_event)
            l = fwd()
            lib.pnce_debug_set(10, 0)
        lo.wait_event(ev)
        fill(lo)
        main.wait_stream(hi); main.wait_stream(lo)
        return l
    def serial():
        hi.wait_stream(main)
        with torch.cuda.stream(hi):
            l = fwd(); fill(hi)
        main.wait_stream(hi)
        return l
    t_z, t_s, t_b = timed(fill), timed(serial), timed(both)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): both()
        torch.cuda.synchronize()
    rows = {e.key: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    ks = '  '.join(f'{k.split("::")[-1][:14]}={v:.0f}' for k, v in rows.items() if 'k_loss' in k or 'k_gather' in k or 'fill' in k)
    print(f'fill {gb:.2f} GiB grid {ctas or "flat"}: fill alone {t_z:.0f} us, serial {t_s:.0f} us, overlapped {t_b:.0f} us (hidden {t_s - t_b:.0f} us)  [{ks}]', flush=True)
