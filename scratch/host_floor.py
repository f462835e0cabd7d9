"""Host floor of the fused call (VERDICT r1 #1b): ms per fwd+bwd step through PatchNCELoss.forward at small batches,
host-side time of the two calls, and the same step replayed from a CUDA graph (the GPU-only time)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
layers = LAYER_SETS["b5"]
out = {}
for B in (1, 8, 16, 24, 32, 64):
    src, tgt = make_maps(layers, B, torch.float32, dev, 1234)
    tgt = [t.requires_grad_() for t in tgt]
    crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])
    torch.manual_seed(7)

    def step():
        for t in tgt:
            t.grad = None
        loss = crit(src, tgt)
        loss.backward()
        return loss

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    n = 300 if B <= 16 else 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hf = hb = 0.0
    e0.record()
    t00 = time.perf_counter()
    for _ in range(n):
        for t in tgt:
            t.grad = None
        t0 = time.perf_counter()
        loss = crit(src, tgt)
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        hf += t1 - t0
        hb += t2 - t1
    host_total = time.perf_counter() - t00
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # the autograd-free entry: forward + backward in one call
    dt = [t.detach() for t in tgt]
    for _ in range(20):
        crit.loss_and_grads(src, dt)
    torch.cuda.synchronize()
    e0.record()
    t00 = time.perf_counter()
    for _ in range(n):
        crit.loss_and_grads(src, dt)
    host_direct = time.perf_counter() - t00
    e1.record()
    torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / n
    # graph replay of the same step (static inputs): GPU-only time
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    try:
        with torch.cuda.graph(g):
            step()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / n
    except Exception as e:  # noqa: BLE001
        gms = repr(e)
    out[B] = {"ms_per_step": round(ms, 4), "host_fwd_us": round(hf / n * 1e6, 1), "host_bwd_us": round(hb / n * 1e6, 1),
              "host_loop_us": round(host_total / n * 1e6, 1),
              "direct_ms": round(dms, 4), "direct_host_us": round(host_direct / n * 1e6, 1), "graph_replay_ms": gms if isinstance(gms, str) else round(gms, 4)}
    print(B, out[B], flush=True)
    if B == 8:
        # GPU timeline of one eager step: kernel start / duration / gap to the previous kernel's end
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        prev_end = None
        for e in evs[-12:]:
            gap = (e.time_range.start - prev_end) if prev_end is not None else 0
            print(f"   {e.name[:60]:60s} dur {e.time_range.elapsed_us():7.1f} us  gap {gap:7.1f} us")
            prev_end = e.time_range.end
    del src, tgt
    torch.cuda.empty_cache()
print(json.dumps(out))
