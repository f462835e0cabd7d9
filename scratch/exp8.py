"""Zero-copy e2e experiment: feature maps stay in pinned HOST memory; the gather kernel reads the sampled
sectors over PCIe through a CUDA alias of the pinned buffer (UVA).  Compared with the bulk H2D copy."""
import sys, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = LAYER_SETS['b5']
shapes = [(B, c, h, w) for c, h, w, _ in layers]

class _Alias:
    def __init__(self, t):
        self.__cuda_array_interface__ = {"shape": tuple(t.shape), "typestr": "<f4", "data": (t.data_ptr(), False),
                                         "version": 2, "strides": None}
def alias(t):
    return torch.as_tensor(_Alias(t), device=dev)

h_src = [torch.randn(s).pin_memory() for s in shapes]
h_tgt = [torch.randn(s).pin_memory() for s in shapes]
crit = pn.PatchNCELoss(0.07, 256)
try:
    a_src = [alias(t) for t in h_src]
    a_tgt = [alias(t).requires_grad_() for t in h_tgt]
    print('alias ok', a_src[0].device, a_src[0].data_ptr() == h_src[0].data_ptr())
except Exception as e:
    print('alias failed', repr(e)); sys.exit(0)
h_loss = torch.empty((), dtype=torch.float32).pin_memory()
def step_zc():
    for t in a_tgt: t.grad = None
    loss = crit(a_src, a_tgt); loss.backward()
    h_loss.copy_(loss.detach(), non_blocking=True)
d_src = [torch.empty(s, device=dev) for s in shapes]
d_tgt = [torch.empty(s, device=dev).requires_grad_() for s in shapes]
def step_bulk():
    for d, h in zip(d_src, h_src): d.copy_(h, non_blocking=True)
    for d, h in zip(d_tgt, h_tgt):
        d.grad = None; d.detach().copy_(h, non_blocking=True)
    loss = crit(d_src, d_tgt); loss.backward()
    h_loss.copy_(loss.detach(), non_blocking=True)
P = sum(min(256, h * w) for _, h, w, _ in layers) * B
for name, fn, n in (('zero-copy', step_zc, 10), ('bulk', step_bulk, 4)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f'{name}: {dt*1e3:.2f} ms/step, {P/dt/1e6:.2f} M patches/s, loss {h_loss.item():.5f}')
# parity of the two paths on the same ids
torch.manual_seed(3); step_zc(); torch.cuda.synchronize(); l1 = h_loss.item(); g1 = a_tgt[1].grad.clone()
torch.manual_seed(3); step_bulk(); torch.cuda.synchronize(); l2 = h_loss.item(); g2 = d_tgt[1].grad
print('same loss', l1, l2, 'grad equal', torch.equal(g1, g2))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step_zc()
    torch.cuda.synchronize()
rows = [(e.key[:70], e.device_time_total / 3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1])[:6]:
    print(f'{t:10.1f} us/step x{c:3d}  {k}')
