"""Kernel breakdown of BASELINE config 4 (B=8, 512^2, P=1024, B5) for NCHW and channels-last maps."""
import sys, json
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import kernel_breakdown, timed_steps
B5_512 = [(64, 512, 512), (256, 128, 128), (256, 128, 128), (128, 256, 256), (64, 512, 512)]
dev = torch.device('cuda', 0)
for layout in ('nchw', 'nhwc'):
    for dtype in (torch.float32, torch.float16):
        g = torch.Generator(device='cuda').manual_seed(1)
        src = [torch.randn(8, *s, device='cuda', generator=g).relu().to(dtype) for s in B5_512]
        tgt = [torch.randn(8, *s, device='cuda', generator=g).relu().to(dtype) for s in B5_512]
        if layout == 'nhwc':
            src = [x.contiguous(memory_format=torch.channels_last) for x in src]
            tgt = [x.contiguous(memory_format=torch.channels_last) for x in tgt]
        tgt = [t.requires_grad_() for t in tgt]
        crit = pn.PatchNCELoss(0.07, 1024)
        def step():
            for t in tgt: t.grad = None
            crit(src, tgt).backward()
        ms = timed_steps(step, 50, 10, 1, dev)
        print(layout, str(dtype), round(ms, 4), json.dumps(kernel_breakdown(step)))
        del src, tgt
