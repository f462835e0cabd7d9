"""Kernel breakdown of the module-split compositions (north-star signatures) at B=64: PatchSampleF(use_mlp=False / True)
+ per-layer PatchNCELoss(feat_q, feat_k)."""
import sys, time
import torch
sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps, kernel_breakdown, timed_steps  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
import os
LIST = os.environ.get("LIST", "1") == "1"
for use_mlp in (False, True):
    samp = pn.PatchSampleF(use_mlp=use_mlp, nc=256).to(dev)
    if use_mlp:
        samp.create_mlp(tgt)
    crit = pn.PatchNCELoss(0.07, 256)
    params = list(samp.parameters())

    def step():
        for t in tgt:
            t.grad = None
        for p_ in params:
            p_.grad = None
        with torch.no_grad():
            feat_k, ids = samp(src, 256, None)
        feat_q, _ = samp(tgt, 256, ids)
        if LIST:
            loss = crit(feat_q, feat_k, batch_size=B)
        else:
            loss = sum(crit(q, k, batch_size=B) for q, k in zip(feat_q, feat_k)) / len(feat_q)
        loss.backward()

    ms = timed_steps(step, 30, 5, 1, dev)
    t0 = time.perf_counter()
    for _ in range(30):
        step()
    host = (time.perf_counter() - t0) / 30 * 1e3
    torch.cuda.synchronize()
    kb = kernel_breakdown(step, 5)
    print(f"use_mlp={use_mlp} B={B}: {ms:.4f} ms/step (host issue {host:.3f} ms), kernels sum {sum(kb.values()):.1f} us: {kb}", flush=True)
    if len(sys.argv) > 2:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(100):
            step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
