"""Why does the loss -> dense programmatic edge cost a loss-kernel's worth of time (PNCE_PDL=31 vs 23)?  Kernel
timeline (start, duration) of a few loss_and_grads steps from the torch profiler."""
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
for B in (8, 64):
    src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
    crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])
    for _ in range(30):
        crit.loss_and_grads(src, tgt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(6):
            crit.loss_and_grads(src, tgt)
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print(f"B={B}")
    prev_end = None
    for e in ev[-12:]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
        print(f"   {e.name[:40]:40s} start {s:9.1f}  dur {d:7.1f}  gap {gap:7.1f}")
        prev_end = e.time_range.end
