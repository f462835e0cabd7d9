"""Round-1 experiments: dense-kernel variants, loss-kernel phase trace, L2 fetch granularity."""
import os, sys, ctypes, time
sys.path.insert(0, '.')
import ctypes
from gan_variant_research_b200 import _lib
lib = _lib.load()
lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
lib.pnce_debug_set_l2_fetch_granularity.restype = ctypes.c_int
gran = os.environ.get('L2GRAN')
if gran:
    print('L2 fetch granularity set before anything else ->', lib.pnce_debug_set_l2_fetch_granularity(int(gran)))
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda')
B = int(os.environ.get('B', '64'))
layers = LAYER_SETS['b5']
src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, math=os.environ.get('MATH', 'tc_bf16x3'))

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def fwd():
    return crit(src, tgt)

print(f'B={B} fwd only: {timeit(fwd):.1f} us')
if gran:
    sys.exit(0)
loss = fwd()
def bwd():
    for t in tgt: t.grad = None
    loss.backward(retain_graph=True)
nbytes = sum(c*h*w for c,h,w,_ in layers) * B * 4
for variant, flags, per_sm, ft in [(0,0,0,128),(0,1,0,128),(0,0,0,64),(0,1,0,64),(2,0,6,0)]:
    lib.pnce_debug_set(4, ft); lib.pnce_debug_set(0, variant); lib.pnce_debug_set(1, flags); lib.pnce_debug_set(2, per_sm)
    us = timeit(bwd)
    print(f'dense variant={variant} flags={flags} ctas/sm={per_sm or "auto"} flat_threads={ft}: {us:.1f} us  {nbytes/us/1e3:.0f} GB/s')
lib.pnce_debug_set(0, 0); lib.pnce_debug_set(1, 0); lib.pnce_debug_set(2, 0)
# correctness of the default (bulk) dense variant vs the warp variant
bwd(); g_tma = [t.grad.clone() for t in tgt]
lib.pnce_debug_set(0, 2); bwd(); g_t2 = [t.grad.clone() for t in tgt]; lib.pnce_debug_set(0, 0)
print('dense flat == bulk bit-exact:', all(torch.equal(a, b) for a, b in zip(g_tma, g_t2)))
lib.pnce_debug_set(0, 1); bwd(); g_warp = [t.grad.clone() for t in tgt]; lib.pnce_debug_set(0, 0)
print('dense tma == warp bit-exact:', all(torch.equal(a, b) for a, b in zip(g_tma, g_warp)))
x = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
print(f'torch zero_: {timeit(lambda: x.zero_()):.1f} us')
# loss-kernel phase trace
tr = torch.zeros(32, dtype=torch.int64, device=dev)
lib.pnce_debug_set(3, tr.data_ptr())
fwd(); torch.cuda.synchronize()
lib.pnce_debug_set(3, 0)
t = tr.cpu().tolist()
names = {0:'epi start',1:'epi prologue done',2:'Z ready',3:'pass A done',4:'pass B done (dz ready)',5:'dQ ready',6:'epi end',
         8:'mma start',9:'mma phase1 issued',10:'mma saw dzready',11:'mma phase2 issued'}
for base, nm in ((0,'CTA 0'),(16,'CTA mid')):
    t0 = t[base]
    print(nm, ', '.join(f'{names[k]}={t[base+k]-t0}' for k in sorted(names)))
