"""CTA timeline of k_loss_tc at BASELINE config 4 (B=8, 512^2, P=1024): start / end of every CTA (= item), per layer.
Needs the experiment build (PNCE_EXPERIMENTS=1)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch, numpy as np
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import make_maps
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
layers = [(64, 512, 512, True), (256, 128, 128, False), (256, 128, 128, False), (128, 256, 256, True), (64, 512, 512, True)]
src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
crit = pn.PatchNCELoss(0.07, 1024)
for _ in range(3): crit(src, tgt)
G = B * 5 * 8
tr = torch.zeros(64 + 3 * G + 1024, dtype=torch.int64, device=dev)
lib.pnce_debug_set(3, tr.data_ptr()); crit(src, tgt); torch.cuda.synchronize(); lib.pnce_debug_set(3, 0)
t = tr.cpu().numpy()
tl = t[64:64 + 3 * G].reshape(G, 3)
base = tl[:, 1].min()
t0 = (tl[:, 1] - base) / 1e3; t1 = (tl[:, 2] - base) / 1e3
print(f'{G} CTAs: start {t0.min():.1f}..{t0.max():.1f} us, end {t1.min():.1f}..{t1.max():.1f} us, makespan {t1.max():.1f}')
dur = t1 - t0
# launch order = heavy first: C=256 (2 layers), C=128, C=64 (2 layers)
n256, n128 = 2 * B * 8, B * 8
for nm, sl in (('C=256', slice(0, n256)), ('C=128', slice(n256, n256 + n128)), ('C=64', slice(n256 + n128, G))):
    d = dur[sl]
    print(f'{nm}: {len(d)} items, duration mean {d.mean():.1f} min {d.min():.1f} max {d.max():.1f} us; starts {t0[sl].min():.1f}..{t0[sl].max():.1f}; ends {t1[sl].min():.1f}..{t1[sl].max():.1f}')
sm = tl[:, 0]
busy = {}
for i in range(G):
    busy.setdefault(int(sm[i]), []).append((t0[i], t1[i]))
ends = sorted(max(e for _, e in v) for v in busy.values())
print(f'{len(busy)} SMs used; per-SM last end: min {ends[0]:.1f} p25 {ends[len(ends)//4]:.1f} p50 {ends[len(ends)//2]:.1f} p75 {ends[3*len(ends)//4]:.1f} max {ends[-1]:.1f}')
print('total busy us / SM-avg:', sum(dur) / 148)
cnt = sorted(len(v) for v in busy.values())
print('items per SM:', {c: cnt.count(c) for c in sorted(set(cnt))})
