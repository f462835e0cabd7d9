"""EMA.update(): the reference's per-parameter loop (ATen ops on the GPU) vs the one-launch multi-tensor kernel,
on the reference generator's 48 parameter tensors (11.4 M values)."""
import sys, json, time
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
shapes = json.load(open('tests/golden/model_param_shapes.json'))['generator']
class Bag(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, device='cuda')) for s in shapes])
net = Bag()
class RefEMA:                                   # restatement of utils/io_ckpt.py:16-29 for timing only
    def __init__(self, model, decay=0.999):
        self.model, self.decay = model, decay
        self.shadow = {n: p.data.clone() for n, p in model.named_parameters() if p.requires_grad}
    def update(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                new = (1.0 - self.decay) * p.data + self.decay * self.shadow[n]
                self.shadow[n] = new.clone()
for name, ema in (('reference loop (ATen, GPU)', RefEMA(net)), ('libpnce one launch', pn.EMA(net))):
    for _ in range(5): ema.update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): ema.update()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    n = sum(p.numel() for p in net.parameters())
    us = e0.elapsed_time(e1) / 200 * 1e3
    print(f'{name}: {us:.1f} us per update (device), {(t1-t0)/200*1e6:.1f} us wall; {n*12/us/1e3:.0f} GB/s of the 12 B/param')
