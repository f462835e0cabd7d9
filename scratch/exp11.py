"""Gather / loss / dense kernel times while S SMs are owned by a spinning blocker kernel."""
import sys, ctypes, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
lib = _lib.load()
lib.pnce_debug_block_sms.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize()
flag = torch.zeros(2, dtype=torch.int32).pin_memory()
side = torch.cuda.Stream()
for S in (0, 16, 32, 40, 48, 64):
    flag.zero_()
    if S:
        lib.pnce_debug_block_sms(S, flag.data_ptr(), side.cuda_stream)
        t0 = time.time()
        while int(flag[1]) < S and time.time() - t0 < 2: pass
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    flag[0] = 1
    torch.cuda.synchronize()
    rows = {e.key[:24]: e.device_time_total / e.count for e in prof.key_averages() if e.device_time_total > 0 and 'pnce::k_' in e.key}
    print(f'S={S} resident={int(flag[1])}: ' + '  '.join(f'{k.split("::")[-1][:14]}={v:.1f}us' for k, v in sorted(rows.items()) if 'blocker' not in k))
