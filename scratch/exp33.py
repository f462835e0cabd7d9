"""D-side row: DiffAugment and the hinge losses, fused launches vs an eager torch formulation of the same maths
(forward + backward, 3x256x256 images, policy color + translation + cutout)."""
import sys
sys.path.insert(0, '.')
import torch, torch.nn.functional as F
import gan_variant_research_b200 as pn
from torch.profiler import profile, ProfilerActivity
def eager_aug(x):
    b, c, h, w = x.shape; dev = x.device
    x = x + (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) - 0.5)
    m = x.mean(1, keepdim=True); x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) * 2) + m
    m = x.mean((1, 2, 3), keepdim=True); x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) + 0.5) + m
    sx, sy = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
    tx = torch.randint(-sx, sx + 1, (b, 1, 1), device=dev); ty = torch.randint(-sy, sy + 1, (b, 1, 1), device=dev)
    rows = (torch.arange(h, device=dev).view(1, h, 1) + tx + 1).clamp(0, h + 1)
    cols = (torch.arange(w, device=dev).view(1, 1, w) + ty + 1).clamp(0, w + 1)
    xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)
    x = xp[torch.arange(b, device=dev).view(b, 1, 1), rows, cols].permute(0, 3, 1, 2)
    ch, cw = int(h * 0.5 + 0.5), int(w * 0.5 + 0.5)
    ox = torch.randint(0, h + (1 - ch % 2), (b, 1, 1), device=dev); oy = torch.randint(0, w + (1 - cw % 2), (b, 1, 1), device=dev)
    gx = (torch.arange(ch, device=dev).view(1, ch, 1) + ox - ch // 2).clamp(0, h - 1)
    gy = (torch.arange(cw, device=dev).view(1, 1, cw) + oy - cw // 2).clamp(0, w - 1)
    mask = torch.ones(b, h, w, dtype=x.dtype, device=dev)
    mask[torch.arange(b, device=dev).view(b, 1, 1), gx, gy] = 0
    return x * mask.unsqueeze(1)
def eager_d(real, fake): return sum(0.5 * (torch.relu(1 - r).mean() + torch.relu(1 + f).mean()) for r, f in zip(real, fake)) / len(real)
def timed(fn, n=100):
    for _ in range(10): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def launches(fn):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    return len([e for e in prof.events() if e.device_type.name == 'CUDA'])
aug = pn.DiffAugment(['color', 'translation', 'cutout'])
for b in (1, 16):
    x = (torch.rand(b, 3, 256, 256, device='cuda') * 2 - 1).requires_grad_()
    up = torch.randn(b, 3, 256, 256, device='cuda')
    def run(f):
        x.grad = None; y = f(x); y.backward(up)
    torch.manual_seed(1); run(eager_aug); g_e = x.grad.clone()
    torch.manual_seed(1); run(aug); g_o = x.grad.clone()
    print(f'B={b}: DiffAugment fwd+bwd eager {timed(lambda: run(eager_aug)):.0f} us / {launches(lambda: run(eager_aug))} launches, '
          f'fused {timed(lambda: run(aug)):.0f} us / {launches(lambda: run(aug))} launches; same draws -> grad diff {float((g_e - g_o).abs().max()):.1e}', flush=True)
    real = [torch.randn(b, 1, 30, 30, device='cuda').requires_grad_()]; fake = [torch.randn(b, 1, 30, 30, device='cuda').requires_grad_()]
    def hr(f):
        real[0].grad = None; fake[0].grad = None; f(real, fake).backward()
    print(f'B={b}: D hinge fwd+bwd eager {timed(lambda: hr(eager_d)):.0f} us / {launches(lambda: hr(eager_d))} launches, '
          f'fused {timed(lambda: hr(pn.discriminator_hinge_loss)):.0f} us / {launches(lambda: hr(pn.discriminator_hinge_loss))} launches', flush=True)
