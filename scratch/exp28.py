import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import gan_variant_research_b200 as pn
from standin_generator import StandInGenerator
from oracle import patchnce_oracle as orc
torch.manual_seed(3)
gen = StandInGenerator(ngf=16, n_blocks=3).cuda()
ids = [0, 2, 4, 6, 16]
photos = torch.randn(2, 3, 64, 64, device="cuda")
def rel(x, y): return float((x - y).abs().max()) / float(x.abs().max())
f1 = gen(photos).detach(); f2 = gen(photos).detach()
print('G forward run-to-run', rel(f1, f2))
with torch.no_grad():
    s1 = gen.get_feature_layers(photos, ids); s2 = gen.get_feature_layers(photos, ids)
    t1 = gen.get_feature_layers(f1, ids)
print('src feats run-to-run', [rel(a, b) for a, b in zip(s1, s2)])
# our loss on fixed maps
def ours(math):
    tt = [t.clone().requires_grad_() for t in t1]
    torch.manual_seed(7)
    loss = pn.PatchNCELoss(0.07, 64, math=math)(s1, tt); loss.backward()
    return loss.item(), [t.grad.clone() for t in tt]
for math in (None, 'simt_f32'):
    a = ours(math); b = ours(math)
    print(math, 'loss', a[0], b[0], 'd tgt run-to-run', [rel(x, y) for x, y in zip(a[1], b[1])])
tt = [t.clone().requires_grad_() for t in t1]
torch.manual_seed(7)
lo, _ = orc.patchnce_loss_torch(s1, tt, 0.07, 64); lo.backward()
print('eager port loss', lo.item(), 'ours vs eager d tgt', [rel(y.grad, x) for x, y in zip(a[1], tt)])
# generator backward given fixed upstream gradients
def gback():
    gen.zero_grad()
    x = f1.clone().requires_grad_()
    fe = gen.get_feature_layers(x, ids)
    torch.autograd.backward(fe, [g for g in a[1]])
    return x.grad.clone()
g1 = gback(); g2 = gback()
print('generator backward run-to-run (fixed d tgt)', rel(g1, g2))
# sensitivity: perturb fake by 1e-6 relative
def full(fk):
    x = fk.clone().requires_grad_()
    torch.manual_seed(7)
    loss = pn.PatchNCELoss(0.07, 64)(s1, gen.get_feature_layers(x, ids)); loss.backward()
    return loss.item(), x.grad.clone()
l1, d1 = full(f1); l2, d2 = full(f1); l3, d3 = full(f1 * (1 + 1e-6 * torch.randn_like(f1)))
print('full chain same input', l1, l2, rel(d1, d2), ' 1e-6 perturbed input', l3, rel(d1, d3))
