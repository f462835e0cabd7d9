"""BASELINE configs 2 / 3 on stand-in networks (the reference's models do not travel to the GPU box): the train_step
order of training/train_cutpp.py:206-331 at 256x256 with a ResNet-9-shaped generator and a 70x70-PatchGAN-shaped
discriminator, (a) from stock torch pieces + the oracle's eager port of the loss, (b) from this package's pieces."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, torch.nn as nn
import gan_variant_research_b200 as pn
import test_train_step_gpu as T
from standin_generator import StandInGenerator
T.NCE_LAYERS = [0, 4, 8, 12, 13]
def make_d():
    L = [nn.Conv2d(3, 64, 4, 2, 1), nn.LeakyReLU(0.2)]
    ch = 64
    for s in (2, 2, 1):
        L += [nn.Conv2d(ch, ch * 2, 4, s, 1), nn.InstanceNorm2d(ch * 2), nn.LeakyReLU(0.2)]; ch *= 2
    return nn.Sequential(*L, nn.Conv2d(ch, 1, 4, 1, 1))
def build(ours):
    torch.manual_seed(0)
    gen, dis = StandInGenerator(ngf=64, n_blocks=9).cuda(), make_d().cuda()
    og = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999)); od = torch.optim.Adam(dis.parameters(), lr=2e-4, betas=(0.5, 0.999))
    sc = torch.amp.GradScaler('cuda')
    if ours:
        pn.enable_encoder_feature_reuse(gen, T.NCE_LAYERS)
        return gen, dis, og, od, sc, pn.EMA(gen, 0.999), pn.DiffAugment(['color', 'translation', 'cutout']), \
            {id(og): pn.FusedAdamStep(og, sc, 10.0), id(od): pn.FusedAdamStep(od, sc, 10.0)}
    return gen, dis, og, od, sc, {n: p.detach().clone() for n, p in gen.named_parameters()}, None, None
# the test's train_step uses num_patches=64; patch it to 256 for the bench through a wrapper of the two loss calls
import oracle.patchnce_oracle as orc
_c0, _c1 = pn.compute_patchnce_loss, orc.compute_patchnce_loss_torch
pn.compute_patchnce_loss = lambda g, a, b, nce_layers, temperature, num_patches: _c0(g, a, b, nce_layers, temperature, 256)
orc.compute_patchnce_loss_torch = lambda g, a, b, l, t, p: _c1(g, a, b, l, t, 256)
for b in (1, 16):
    photos = torch.rand(b, 3, 256, 256, device='cuda') * 2 - 1
    res = {}
    for ours in (False, True):
        st = build(ours); torch.manual_seed(5)
        for _ in range(3): T.train_step(st, photos, ours)
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 10
        for _ in range(n): out = T.train_step(st, photos, ours)
        torch.cuda.synchronize(); res[ours] = (time.perf_counter() - t0) / n * 1e3
        del st; torch.cuda.empty_cache()
    print(f'B={b}: train_step (with its four .item() reads) stock pieces {res[False]:.1f} ms, this package {res[True]:.1f} ms ({res[False] / res[True]:.2f}x); last losses {out}', flush=True)
