"""Per-item phase stamps of the persistent loss kernel (CTA 0 and CTA grid/2)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch, numpy as np
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
crit = pn.PatchNCELoss(0.07, 256)
for _ in range(3): crit(src, tgt)
G = 148
tr = torch.zeros(64 + 3 * G + 512 + 64, dtype=torch.int64, device=dev)
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if len(sys.argv) > 3: lib.pnce_debug_set(9, int(sys.argv[3]))
lib.pnce_debug_set(13, rep); lib.pnce_debug_set(3, tr.data_ptr()); crit(src, tgt); torch.cuda.synchronize(); lib.pnce_debug_set(3, 0); lib.pnce_debug_set(13, 0)
t = tr.cpu().numpy()
tl = t[64:64 + 3 * G].reshape(G, 3)
t0 = (tl[:, 1] - tl[:, 1].min()) / 1e3; t1 = (tl[:, 2] - tl[:, 1].min()) / 1e3
print(f'CTA timeline: start {t0.min():.1f}..{t0.max():.1f} us, end {t1.min():.1f}..{t1.max():.1f} us, makespan {t1.max():.1f}')
names = {0: 'E:norm', 1: 'E:zfull', 2: 'E:A', 3: 'E:B', 4: 'E:dqfull', 5: 'E:dQ', 6: 'M:zfree', 7: 'M:P1', 8: 'M:dz0', 9: 'M:P2', 10: 'P:ring', 11: 'P:P1ld', 12: 'P:zfull', 13: 'P:P2ld', 14: 'E:preB', 15: 'E:B0ld', 16: 'q:in', 17: 'q:ld0', 18: 'q:qa', 19: 'q:kc', 20: 'q:ch0', 21: 'q:iss', 22: 'q:ld1', 23: 'q:qb', 24: 'q:ch1', 25: 'q:end0', 26: 'q:ld2', 27: 'q:ch2', 28: 'q:ld3', 29: 'q:ch3', 30: 'q:qa2', 31: 'q:kc2'}
for sel, nm in ((0, 'CTA 0'), (1, 'CTA grid/2')):
    st = t[64 + 3 * G + sel * 256: 64 + 3 * G + sel * 256 + 256].reshape(8, 32)
    ref = st[st > 0].min()
    print(nm)
    for n in range(8):
        if st[n].max() == 0: continue
        print(f'  item {n}: ' + '  '.join(f'{names[k]}={int(st[n, k] - ref)}' for k in sorted(names, key=lambda k: st[n, k]) if st[n, k] > 0))
