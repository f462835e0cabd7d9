"""Kernel breakdown (torch profiler) of the fused head step at small batches.  python scratch/head_small_breakdown.py [B ...]"""
import sys
import torch
sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps, kernel_breakdown  # noqa: E402

dev = torch.device("cuda", 0)
for B in [int(a) for a in sys.argv[1:]] or [1, 16]:
    src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
    netF = pn.PatchSampleF(use_mlp=True, nc=256).to(dev)
    netF.create_mlp(tgt)
    params = list(netF.parameters())

    def dstep():
        for p_ in params:
            p_.grad = None
        pn.head_loss_and_grads(netF, src, tgt, 0.07, 256)

    for _ in range(10):
        dstep()
    kb = kernel_breakdown(dstep, 20)
    print(B, sum(kb.values()), kb, flush=True)
