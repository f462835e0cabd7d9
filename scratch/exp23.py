"""Head-mode gradient mismatch found by stress2: which kernel?  Same case with the persistent kernels on / off."""
import sys, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from oracle import patchnce_oracle as orc
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
def relerr(got, want):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-12))
def run(b, shapes, p, nc, seed, math='tc_bf16x3'):
    g = torch.Generator().manual_seed(seed)
    nl = len(shapes)
    src = [torch.randn(b, *s, generator=g) for s in shapes]; tgt = [torch.randn(b, *s, generator=g) for s in shapes]
    ids = [torch.randint(0, s[1] * s[2], (min(p, s[1] * s[2]),), generator=g) for s in shapes]
    torch.manual_seed(seed)
    netF = pn.PatchSampleF(use_mlp=True, nc=nc, init_gain=0.3); netF.create_mlp([x.cuda() for x in tgt])
    for prm in netF.parameters():
        if prm.dim() == 1: torch.nn.init.normal_(prm, 0.0, 0.1)
    heads = [tuple(x.detach().cpu().clone().requires_grad_() for x in (m[0].weight, m[0].bias, m[2].weight, m[2].bias))
             for m in (getattr(netF, f'mlp_{l}') for l in range(nl))]
    tc = [x.clone().requires_grad_() for x in tgt]
    want = orc.patchnce_head_loss_torch(src, tc, ids, heads); want.backward()
    out = {}
    for knob in (0, 1):
        lib.pnce_debug_set(6, knob)
        netF.zero_grad()
        t = [x.cuda().requires_grad_() for x in tgt]
        loss, _ = pn.patchnce_with_head(netF, [x.cuda() for x in src], t, 0.07, p, [i.cuda() for i in ids], fused=True, math=math)
        loss.backward()
        e_t = max(relerr(t[l].grad.cpu().numpy(), tc[l].grad.numpy()) for l in range(nl))
        e_w = [[relerr(gp.grad.cpu().numpy(), w.grad.numpy()) for gp, w in zip((m[0].weight, m[0].bias, m[2].weight, m[2].bias), heads[l])]
               for l, m in enumerate(getattr(netF, f'mlp_{l}') for l in range(nl))]
        out[knob] = (abs(loss.item() - want.item()) / abs(want.item()), e_t, e_w)
    lib.pnce_debug_set(6, 0)
    # also the module-split (non-fused) composition as a third opinion
    netF.zero_grad()
    t = [x.cuda().requires_grad_() for x in tgt]
    loss, _ = pn.patchnce_with_head(netF, [x.cuda() for x in src], t, 0.07, p, [i.cuda() for i in ids], fused=False)
    loss.backward()
    e_t = max(relerr(t[l].grad.cpu().numpy(), tc[l].grad.numpy()) for l in range(nl))
    e_w = [[relerr(gp.grad.cpu().numpy(), w.grad.numpy()) for gp, w in zip((m[0].weight, m[0].bias, m[2].weight, m[2].bias), heads[l])]
           for l, m in enumerate(getattr(netF, f'mlp_{l}') for l in range(nl))]
    print(f'b={b} shapes={shapes} p={p} nc={nc} math={math}')
    for knob in (0, 1):
        le, et, ew = out[knob]
        print(f'   persistent={"on" if knob == 0 else "off"}: loss {le:.1e} d tgt {et:.1e} dW/db per layer', [[f'{x:.0e}' for x in r] for r in ew])
    print(f'   module split (ATen Linear): d tgt {e_t:.1e} dW', [[f'{x:.0e}' for x in r] for r in e_w])
run(1, [(128, 14, 9), (24, 11, 24)], 128, 256, 3)
run(1, [(32, 14, 19), (32, 13, 21)], 256, 256, 4)
run(3, [(64, 32, 32), (128, 16, 16), (24, 20, 12)], 64, 256, 123)
run(1, [(32, 14, 19)], 256, 256, 5)
run(1, [(32, 14, 19)], 256, 256, 5, math='tc_bf16')
run(4, [(32, 14, 19)], 256, 256, 5)
