"""Head backward defect: error pattern of db1 / dW1 over hidden units for the failing case."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import gan_variant_research_b200 as pn
from oracle import patchnce_oracle as orc
def run(b, shapes, p, nc, seed, gseed):
    g = torch.Generator().manual_seed(gseed); nl = len(shapes)
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
    t = [x.cuda().requires_grad_() for x in tgt]
    loss, _ = pn.patchnce_with_head(netF, [x.cuda() for x in src], t, 0.07, p, [i.cuda() for i in ids], fused=True)
    loss.backward()
    print(f'b={b} shapes={shapes} p={p} nc={nc}')
    for l in range(nl):
        m = getattr(netF, f'mlp_{l}')
        db1 = m[0].bias.grad.cpu().numpy(); wb1 = heads[l][1].grad.numpy()
        e = np.abs(db1 - wb1) / np.abs(wb1).max()
        bad = np.nonzero(e > 1e-4)[0]
        dW1 = m[0].weight.grad.cpu().numpy(); wW1 = heads[l][0].grad.numpy()
        eW = np.abs(dW1 - wW1).max(axis=1) / np.abs(wW1).max()
        badW = np.nonzero(eW > 1e-4)[0]
        dt = t[l].grad.cpu().numpy(); wt = tc[l].grad.numpy()
        et = np.abs(dt - wt).reshape(b, -1).max(axis=1) / np.abs(wt).max()
        print(f'  layer {l}: db1 bad units {bad.tolist()[:20]} (n={len(bad)}) err {e[bad][:6]}; dW1 bad rows {badW.tolist()[:20]} (n={len(badW)}); d tgt err per image {np.round(et, 5).tolist()}')
        if len(bad):
            im = int(np.argmax(et))
            err_map = np.abs(dt - wt)[im].max(axis=0).reshape(-1)
            pos = np.argsort(-err_map)[:4]
            idl = ids[l].numpy()
            print(f'    image {im}: worst positions {pos.tolist()} errs {err_map[pos]}; rows sampling them:', [np.nonzero(idl == q)[0].tolist() for q in pos])
            srt = np.sort(idl); 
            print(f'    sorted slot(s) of the worst position: {np.nonzero(srt == pos[0])[0].tolist()}  (P={len(idl)})')
            j = bad[0]
            print(f'    unit {j}: got db1 {db1[j]:.6e} want {wb1[j]:.6e} diff {db1[j]-wb1[j]:.3e}')
run(3, [(100, 23, 9), (64, 13, 23)], 128, 128, 102, 10102)
run(2, [(64, 14, 4), (128, 8, 4)], 64, 128, 104, 10104)
