"""Encoder-feature reuse, timed on a ResNet-9-shaped stand-in generator (tests/standin_generator.py, ngf=64, 9 blocks):
fake = G(photos); compute_patchnce_loss(G, photos, fake); backward -- with and without the tap, fp32 and autocast."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import gan_variant_research_b200 as pn
from standin_generator import StandInGenerator
from oracle import patchnce_oracle as orc
torch.manual_seed(0)
gen = StandInGenerator(ngf=64, n_blocks=9).cuda()
ids = [0, 4, 8, 12, 13]
def step(photos, reuse, amp, eager=False):
    gen.zero_grad(set_to_none=True)
    with torch.autocast('cuda', enabled=amp):
        fake = gen(photos)
        if eager:
            with torch.no_grad(): s = gen.get_feature_layers(photos, ids)
            loss, _ = orc.patchnce_loss_torch([f.detach() for f in s], gen.get_feature_layers(fake, ids), 0.07, 256)
        else:
            loss = pn.compute_patchnce_loss(gen, photos, fake, ids, 0.07, 256)
    loss.backward()
    return loss
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for b in (1, 4, 16):
    photos = torch.rand(b, 3, 256, 256, device='cuda') * 2 - 1
    for amp in (False, True):
        t_e = timed(lambda: step(photos, False, amp, eager=True), 5)
        t0 = timed(lambda: step(photos, False, amp))
        cache = pn.enable_encoder_feature_reuse(gen, ids)
        t1 = timed(lambda: step(photos, True, amp))
        assert cache.hits >= 13 and cache.misses == 0, (cache.hits, cache.misses)
        cache.remove()
        torch.manual_seed(7); l0 = step(photos, False, amp).item()
        cache = pn.enable_encoder_feature_reuse(gen, ids)
        torch.manual_seed(7); l1 = step(photos, True, amp).item(); cache.remove()
        print(f'B={b} amp={amp}: G fwd + PatchNCE + backward: eager reference port {t_e:.2f} ms | this path {t0:.2f} ms | + feature reuse {t1:.2f} ms '
              f'(saves {t0 - t1:.2f} ms = {100 * (t0 - t1) / t0:.0f}%); loss {l0:.6f} vs {l1:.6f}', flush=True)
