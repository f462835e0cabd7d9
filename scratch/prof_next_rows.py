"""One pass over the kernels of the "next" rows for ncu: EMA, AMP optimiser step on a ResNet-9-sized
parameter set, DiffAugment forward/backward at B=16, hinge losses."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import gan_variant_research_b200 as pn
from standin_generator import StandInGenerator
torch.manual_seed(0)
gen = StandInGenerator(ngf=64, n_blocks=9).cuda()
opt = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
sc = torch.amp.GradScaler('cuda', init_scale=1.0, growth_interval=10 ** 9); sc.scale(torch.zeros((), device='cuda'))
for p in gen.parameters(): p.grad = torch.randn_like(p) * 1e-4
st = pn.FusedAdamStep(opt, sc, 10.0); ema = pn.EMA(gen, 0.999)
aug = pn.DiffAugment(['color', 'translation', 'cutout'])
x = (torch.rand(16, 3, 256, 256, device='cuda') * 2 - 1).requires_grad_(); up = torch.randn_like(x)
r = torch.randn(16, 1, 30, 30, device='cuda').requires_grad_(); f = torch.randn(16, 1, 30, 30, device='cuda').requires_grad_()
for _ in range(3):
    st.step(); ema.update()
    x.grad = None; aug(x).backward(up)
    r.grad = None; f.grad = None; pn.discriminator_hinge_loss(r, f).backward()
torch.cuda.synchronize()
print('ok', float(st.last_total_norm()))
