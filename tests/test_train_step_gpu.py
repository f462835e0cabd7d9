"""Every piece of this package in the order the reference's train_step uses them (training/train_cutpp.py:206-331):
D step (generator forward, DiffAugment on both inputs, hinge loss, scaled backward, AMP optimiser step), G step
(generator forward, DiffAugment, adversarial hinge, PatchNCE with encoder-feature reuse, scaled backward, AMP optimiser
step), EMA update -- on stand-in networks with the reference's layout, under autocast, against the same loop built from
stock torch pieces and the oracle's eager port of the loss.  Same seeds: the first step's losses must agree (later
steps diverge chaotically on randomly initialised networks: a 1e-6 change of `fake` moves the gradient by percent)."""
import pytest
import torch
import torch.nn as nn

from standin_generator import StandInGenerator

pytestmark = pytest.mark.gpu
NCE_LAYERS = [0, 2, 4, 6, 16]


def make_discriminator():
    return nn.Sequential(nn.Conv2d(3, 16, 4, 2, 1), nn.LeakyReLU(0.2), nn.Conv2d(16, 32, 4, 2, 1), nn.InstanceNorm2d(32),
                         nn.LeakyReLU(0.2), nn.Conv2d(32, 1, 4, 1, 1))


def build(ours):
    import gan_variant_research_b200 as pn
    torch.manual_seed(0)
    gen, dis = StandInGenerator(ngf=16, n_blocks=3).cuda(), make_discriminator().cuda()
    opt_g = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_d = torch.optim.Adam(dis.parameters(), lr=2e-4, betas=(0.5, 0.999))
    scaler = torch.amp.GradScaler("cuda", init_scale=256.0)
    if ours:
        pn.enable_encoder_feature_reuse(gen, NCE_LAYERS)
        ema = pn.EMA(gen, 0.999)
        aug = pn.DiffAugment(["color", "translation", "cutout"])
        steppers = {id(opt_g): pn.FusedAdamStep(opt_g, scaler, 10.0), id(opt_d): pn.FusedAdamStep(opt_d, scaler, 10.0)}
    else:
        ema = {n: p.detach().clone() for n, p in gen.named_parameters()}
        aug = None
        steppers = None
    return gen, dis, opt_g, opt_d, scaler, ema, aug, steppers


def eager_aug(x):
    """The maths of DiffAugment(['color', 'translation', 'cutout']) in plain torch, same draws in the same order."""
    import torch.nn.functional as F
    b, c, h, w = x.shape
    dev = x.device
    x = x + (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) - 0.5)
    m = x.mean(1, keepdim=True)
    x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) * 2) + m
    m = x.mean((1, 2, 3), keepdim=True)
    x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) + 0.5) + m
    sx, sy = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
    tx = torch.randint(-sx, sx + 1, (b, 1, 1), device=dev)
    ty = torch.randint(-sy, sy + 1, (b, 1, 1), device=dev)
    rows = (torch.arange(h, device=dev).view(1, h, 1) + tx + 1).clamp(0, h + 1)
    cols = (torch.arange(w, device=dev).view(1, 1, w) + ty + 1).clamp(0, w + 1)
    xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)
    x = xp[torch.arange(b, device=dev).view(b, 1, 1), rows, cols].permute(0, 3, 1, 2)
    ch, cw = int(h * 0.5 + 0.5), int(w * 0.5 + 0.5)
    ox = torch.randint(0, h + (1 - ch % 2), (b, 1, 1), device=dev)
    oy = torch.randint(0, w + (1 - cw % 2), (b, 1, 1), device=dev)
    gx = (torch.arange(ch, device=dev).view(1, ch, 1) + ox - ch // 2).clamp(0, h - 1)
    gy = (torch.arange(cw, device=dev).view(1, 1, cw) + oy - cw // 2).clamp(0, w - 1)
    mask = torch.ones(b, h, w, dtype=x.dtype, device=dev)
    mask[torch.arange(b, device=dev).view(b, 1, 1), gx, gy] = 0
    return x * mask.unsqueeze(1)


def train_step(state, photos, ours):
    import gan_variant_research_b200 as pn
    from oracle import patchnce_oracle as orc
    gen, dis, opt_g, opt_d, scaler, ema, aug, steppers = state
    augment = aug if ours else eager_aug

    def optimiser_step(opt):                                          # AMPContext.step_optimizer, amp_utils.py:29-41
        if ours:
            steppers[id(opt)].step()
        else:
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_([p for g in opt.param_groups for p in g["params"] if p.grad is not None], 10.0)
            scaler.step(opt)
            scaler.update()

    # ---- D step (train_cutpp.py:229-253)
    opt_d.zero_grad()
    with torch.autocast("cuda"):
        fake = gen(photos)
        real_pred, fake_pred = dis(augment(photos)), dis(augment(fake.detach()))
        if ours:
            d_loss = pn.discriminator_hinge_loss(real_pred, fake_pred)
        else:
            d_loss = 0.5 * (torch.relu(1.0 - real_pred).mean() + torch.relu(1.0 + fake_pred).mean())
    scaler.scale(d_loss).backward()
    optimiser_step(opt_d)
    # ---- G step (:266-308)
    opt_g.zero_grad()
    with torch.autocast("cuda"):
        fake = gen(photos)
        pred = dis(augment(fake))
        if ours:
            g_adv = pn.generator_hinge_loss(pred)
            nce = pn.compute_patchnce_loss(gen, photos, fake, nce_layers=NCE_LAYERS, temperature=0.07, num_patches=64)
        else:
            g_adv = -pred.mean()
            nce = orc.compute_patchnce_loss_torch(gen, photos, fake, NCE_LAYERS, 0.07, 64)
        g_loss = 1.0 * g_adv + 1.0 * nce
    scaler.scale(g_loss).backward()
    optimiser_step(opt_g)
    # ---- EMA (:310-312)
    if ours:
        ema.update()
    else:
        with torch.no_grad():
            for n, p in gen.named_parameters():
                ema[n] = (1.0 - 0.999) * p + 0.999 * ema[n]
    return {"d_loss": d_loss.item(), "g_adv": g_adv.item(), "nce": nce.item(), "scale": scaler.get_scale()}


def test_train_step_with_every_piece_matches_the_stock_loop():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    photos = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    logs = {}
    for ours in (False, True):
        state = build(ours)
        torch.manual_seed(99)
        logs[ours] = [train_step(state, photos, ours) for _ in range(3)]
        gen, ema = state[0], state[5]
        if ours:
            cache = getattr(gen, "_pnce_encoder_feature_cache")
            assert cache.hits == 3                                    # one generator pass saved per G step
            shadow = ema.shadow
            assert pn.poll_nonfinite_warnings(block=True) == 0
        else:
            shadow = ema
        for n, p in gen.named_parameters():
            assert torch.isfinite(p).all() and torch.isfinite(shadow[n]).all()
            if p.dim() > 1:
                assert not torch.equal(shadow[n], p)                  # parameters moved, the EMA lags behind
    a, b = logs[False][0], logs[True][0]
    # autocast: the stock loop runs the loss in fp16 GEMMs (torch.mm under autocast), ours in bf16x3 on fp16 maps
    assert b["d_loss"] == pytest.approx(a["d_loss"], rel=2e-3)
    assert b["g_adv"] == pytest.approx(a["g_adv"], rel=5e-3, abs=2e-3)
    assert b["nce"] == pytest.approx(a["nce"], rel=5e-3)
    for rec in logs[True] + logs[False]:
        assert all(v == v and abs(v) < 1e4 for v in rec.values())
    assert [r["scale"] for r in logs[True]] == [r["scale"] for r in logs[False]]
