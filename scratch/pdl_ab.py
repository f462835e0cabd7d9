"""A/B of programmatic dependent launch by launch site (PNCE_PDL mask: 1 prep, 2 gather, 4 loss, 8 dense, 16 rest):
ms per step of the autograd route and of loss_and_grads at small and large batches."""
import sys
import torch
sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
res = {}
for B in (1, 8, 16, 64):
    src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
    tgt = [t.requires_grad_() for t in tgt]
    crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])
    dt = [t.detach() for t in tgt]

    def eager():
        for t in tgt:
            t.grad = None
        crit(src, tgt).backward()

    def direct():
        crit.loss_and_grads(src, dt)

    row = []
    for fn in (eager, direct):
        for _ in range(30):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            n = 300 if B <= 16 else 100
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        row.append(round(best * 1e3, 1))
    res[B] = row
print(res)
