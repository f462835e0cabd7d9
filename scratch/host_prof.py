"""Where does the host time of a small-batch step go?  cProfile over fwd+bwd at B=1 plus raw timings of the C calls."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from gan_variant_research_b200 import _lib, patchnce as pm  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])
torch.manual_seed(7)


def step():
    for t in tgt:
        t.grad = None
    loss = crit(src, tgt)
    loss.backward()


for _ in range(50):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

# raw C calls, GPU kept drained so the launch queue never backs up
lib = _lib.load()
sp = pm._shape_plan(tgt, [256] * 5, "tc_bf16x3")
ws = torch.empty(sp.ws_bytes, dtype=torch.uint8, device=dev)
out = torch.empty(6, dtype=torch.float32, device=dev)
ids_all = torch.empty(5 * 256, dtype=torch.int64, device=dev)
grads = [torch.empty_like(t) for t in tgt]
L = sp.fwd_layers
for l in range(5):
    L[l].src, L[l].tgt, L[l].ids, L[l].dtgt = src[l].data_ptr(), tgt[l].data_ptr(), ids_all.data_ptr() + 2048 * l, grads[l].data_ptr()
st = torch.cuda.current_stream().cuda_stream
one = torch.ones((), device=dev)
for name, fn in (("pnce_fwd_draw", lambda: lib.pnce_fwd_draw(L, 5, B, 0, 0.07, 1, ws.data_ptr(), sp.ws_bytes, 7, 0, out.data_ptr(), None, st)),
                 ("pnce_bwd", lambda: lib.pnce_bwd(L, 5, B, 0, 1, ws.data_ptr(), sp.ws_bytes, one.data_ptr(), st))):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        rc = fn()
        ts.append(time.perf_counter() - t0)
        assert rc == 0
        if len(ts) % 8 == 0:
            torch.cuda.synchronize()
    ts.sort()
    print(f"{name}: host p50 {ts[100] * 1e6:.1f} us  p10 {ts[20] * 1e6:.1f}  p90 {ts[180] * 1e6:.1f}")
for name, fn in (("torch.empty ws", lambda: torch.empty(sp.ws_bytes, dtype=torch.uint8, device=dev)),
                 ("empty_like x5", lambda: [torch.empty_like(t) for t in tgt]),
                 ("current_stream", lambda: torch.cuda.current_stream(dev).cuda_stream),
                 ("is_capturing", torch.cuda.is_current_stream_capturing),
                 ("philox_take", lambda: pm._philox_take(dev, 5)),
                 ("grad=None x5", lambda: [setattr(t, "grad", None) for t in tgt])):
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    print(f"{name}: {(time.perf_counter() - t0) / 2000 * 1e6:.2f} us")
