"""Feasibility: zero fill of 2.15 GB (the two 64 x 256^2 gradient maps at B=64) on a SIDE stream while the forward's gather
runs -- on plain and on compressible memory (where the fill needs a quarter of the DRAM bandwidth).  How much longer does
forward + fill take than forward alone?"""
import sys
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import patchnce as pm
from bench import LAYER_SETS, make_maps, kernel_breakdown
dev = torch.device('cuda', 0); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
crit = pn.PatchNCELoss(0.07, 256)
nbytes = 2 * B * 64 * 256 * 256 * 4
plain = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
with torch.cuda.use_mem_pool(pm._grad_pool(dev), dev):
    comp = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
print('compressed:', pn.gradient_is_compressed(comp))
side = torch.cuda.Stream()
def fwd_only():
    call, t, _ = crit._begin(src, tgt)
    return pm._run_fwd(call, t)
def run(buf, n=30):
    main = torch.cuda.current_stream()
    for _ in range(3):
        fwd_only()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        if buf is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                buf.zero_()
        fwd_only()
        if buf is not None:
            main.wait_stream(side)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def fill_only(buf, n=30):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        buf.zero_()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f'forward alone {run(None):.1f} us')
print(f'fill alone: plain {fill_only(plain):.1f} us, compressible {fill_only(comp):.1f} us')
print(f'forward || fill(plain) {run(plain):.1f} us')
print(f'forward || fill(compressible) {run(comp):.1f} us')
