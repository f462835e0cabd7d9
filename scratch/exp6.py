"""Timeline of k_loss_tc CTAs (debug trace): per-CTA durations by layer, per-SM occupancy, makespan."""
import sys, ctypes
sys.path.insert(0, '.')
import torch, numpy as np
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = 64
layers = LAYER_SETS['b5']
src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
crit = pn.PatchNCELoss(0.07, 256)
for _ in range(3): crit(src, tgt)
n = B * 10
tr = torch.zeros(64 + 3 * n, dtype=torch.int64, device=dev)
lib.pnce_debug_set(3, tr.data_ptr()); crit(src, tgt); torch.cuda.synchronize(); lib.pnce_debug_set(3, 0)
t = tr.cpu().numpy()[64:].reshape(n, 3)
sm, t0, t1 = t[:, 0], t[:, 1], t[:, 2]
base = t0.min(); t0 = (t0 - base) / 1e3; t1 = (t1 - base) / 1e3
dur = t1 - t0
print(f'makespan {t1.max():.1f} us, CTAs {n}, sum of durations / 148 SMs = {dur.sum()/148:.1f} us')
# heavy-first order: C = 256,256,128,64,64 -> 128 CTAs each
for k, name in enumerate(['C=256 (a)', 'C=256 (b)', 'C=128', 'C=64 (a)', 'C=64 (b)']):
    d = dur[k*128:(k+1)*128]; s = t0[k*128:(k+1)*128]
    print(f'{name}: duration mean {d.mean():.1f} min {d.min():.1f} max {d.max():.1f} us; start {s.min():.1f}..{s.max():.1f} us')
busy = np.zeros(148)
for s_, d_ in zip(sm, dur): busy[int(s_)] += d_
print(f'per-SM busy: mean {busy.mean():.1f} min {busy.min():.1f} max {busy.max():.1f} us; SMs used {len(set(sm.tolist()))}')
order = np.argsort(t0)
print('first 10 starts', np.round(t0[order][:10], 1), 'last 5 ends', np.round(np.sort(t1)[-5:], 1))

# per-phase clock64 stamps of CTA 0 (C=256 layer) and CTA grid/2: epilogue slots 0..6, MMA thread 8..11
raw = tr.cpu().numpy()[:32]
for base_, name in ((0, 'CTA 0'), (16, 'CTA grid/2')):
    st = raw[base_:base_ + 16]; ref = min(x for x in st if x > 0)
    print(name, 'cycles since first stamp:', {k: int(v - ref) for k, v in enumerate(st) if v > 0})
