import sys, time, json
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
B5_512 = [(64, 512, 512), (256, 128, 128), (256, 128, 128), (128, 256, 256), (64, 512, 512)]
for layout in ('nchw', 'nhwc'):
    g = torch.Generator(device='cuda').manual_seed(1)
    src = [torch.randn(8, *s, device='cuda', generator=g).relu() for s in B5_512]
    tgt = [torch.randn(8, *s, device='cuda', generator=g).relu() for s in B5_512]
    if layout == 'nhwc':
        src = [x.contiguous(memory_format=torch.channels_last) for x in src]
        tgt = [x.contiguous(memory_format=torch.channels_last) for x in tgt]
    tgt = [t.requires_grad_() for t in tgt]
    crit = pn.PatchNCELoss(0.07, 1024)
    for _ in range(10):
        for t in tgt: t.grad = None
        crit(src, tgt).backward()
    torch.cuda.synchronize()
    hf = hb = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 50
    for _ in range(n):
        for t in tgt: t.grad = None
        t0 = time.perf_counter(); loss = crit(src, tgt); t1 = time.perf_counter(); loss.backward(); t2 = time.perf_counter()
        hf += t1 - t0; hb += t2 - t1
    e1.record(); torch.cuda.synchronize()
    print(layout, 'ms/step', round(e0.elapsed_time(e1) / n, 4), 'host fwd us', round(hf / n * 1e6, 1), 'host bwd us', round(hb / n * 1e6, 1))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            for t in tgt: t.grad = None
            crit(src, tgt).backward()
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    prev = None
    for e in evs[-10:]:
        gap = (e.time_range.start - prev) if prev is not None else 0
        print(f"   {e.name[:50]:50s} dur {e.time_range.elapsed_us():7.1f} gap {gap:7.1f}")
        prev = e.time_range.end
    del src, tgt
