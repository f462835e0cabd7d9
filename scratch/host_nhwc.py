import sys, time
sys.path.insert(0, '.')
import torch, gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
dev = torch.device('cuda', 0)
for layout in ('nchw', 'nhwc'):
    src, tgt = make_maps(LAYER_SETS['b5'], 8, torch.float32, dev, 1)
    if layout == 'nhwc':
        src = [x.contiguous(memory_format=torch.channels_last) for x in src]
        tgt = [x.contiguous(memory_format=torch.channels_last) for x in tgt]
    tgt = [t.requires_grad_() for t in tgt]
    crit = pn.PatchNCELoss(0.07, 256)
    def step():
        for t in tgt: t.grad = None
        crit(src, tgt).backward()
    for _ in range(30): step()
    torch.cuda.synchronize()
    hf = hb = 0.0; n = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        for t in tgt: t.grad = None
        t0 = time.perf_counter(); loss = crit(src, tgt); t1 = time.perf_counter(); loss.backward(); t2 = time.perf_counter()
        hf += t1 - t0; hb += t2 - t1
    e1.record(); torch.cuda.synchronize()
    print(layout, 'ms', round(e0.elapsed_time(e1) / n, 4), 'host fwd', round(hf / n * 1e6, 1), 'bwd', round(hb / n * 1e6, 1))
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): step()
    pr.disable(); torch.cuda.synchronize()
    st = pstats.Stats(pr); st.sort_stats('tottime'); st.print_stats(12)
