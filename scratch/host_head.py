"""Head mode (patchnce_with_head, nc=256) at small batches: device-synchronised step time, host issue time of forward and
backward, cProfile of the step, and the GPU time of the same step from CUDA-graph-free event timing with a deep queue.
python scratch/host_head.py [B ...]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
batches = [int(a) for a in sys.argv[1:]] or [1, 8, 16, 64]
for B in batches:
    src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
    tgt = [t.requires_grad_() for t in tgt]
    netF = pn.PatchSampleF(use_mlp=True, nc=256).to(dev)
    netF.create_mlp(tgt)
    params = list(netF.parameters())
    torch.manual_seed(7)
    hf, hb = [], []

    def step(rec=False):
        for t in tgt:
            t.grad = None
        for p_ in params:
            p_.grad = None
        t0 = time.perf_counter()
        loss, _ = pn.patchnce_with_head(netF, src, tgt, 0.07, 256)
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        if rec:
            hf.append(t1 - t0); hb.append(t2 - t1)

    for _ in range(30):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        step(True)
    e1.record()
    torch.cuda.synchronize()
    hf.sort(); hb.sort()
    print(f"B={B} head step {e0.elapsed_time(e1) / 200:.4f} ms   host fwd p50 {hf[100] * 1e6:.1f} us  bwd p50 {hb[100] * 1e6:.1f} us", flush=True)
    d_tgt = [t.detach() for t in tgt]
    hd = []

    def dstep(rec=False):
        for p_ in params:
            p_.grad = None
        t0 = time.perf_counter()
        pn.head_loss_and_grads(netF, src, d_tgt, 0.07, 256)
        if rec:
            hd.append(time.perf_counter() - t0)

    for _ in range(30):
        dstep()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        dstep(True)
    e1.record()
    torch.cuda.synchronize()
    hd.sort()
    print(f"B={B} head direct {e0.elapsed_time(e1) / 200:.4f} ms   host p50 {hd[100] * 1e6:.1f} us", flush=True)
    if B == batches[0]:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("tottime").print_stats(30)
    del src, tgt, netF
    torch.cuda.empty_cache()
