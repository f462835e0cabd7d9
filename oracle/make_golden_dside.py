"""Golden fixture for the D-side row from the UNMODIFIED reference:
training/diffaugment.py::DiffAugment and losses/adv_hinge.py on the CPU.  Build container only (needs /root/reference).
The random parameters the reference drew are recovered by re-seeding and replaying the draws in the order the B200 host
makes them (gan_variant_research_b200/dside.py) -- if that order differed from the reference's, the oracle fed with the
replayed parameters would not reproduce the reference's output (tests/test_dside.py checks exactly that)."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PNCE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "GAN_Variant1"))
sys.path.insert(0, REF)
from GAN_Variant1.training.diffaugment import DiffAugment  # noqa: E402
from GAN_Variant1.losses.adv_hinge import discriminator_hinge_loss, generator_hinge_loss  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
POLICIES = [["color", "translation", "cutout"], ["color", "translation", "cutout_light"], ["translation"], ["color"],
            ["cutout"], ["translation", "cutout_light"]]
SHAPE = (3, 3, 20, 24)
CUT = {"cutout": 0.5, "cutout_light": 0.2}


def replay_draws(policy, shape, seed):
    """The draws of DiffAugment(policy)(x) in order: 3 x rand for color, 2 x randint for translation, 2 for cutout."""
    b, _, h, w = shape
    torch.manual_seed(seed)
    out = {}
    if "color" in policy:
        out["color"] = [torch.rand(b, 1, 1, 1).numpy().reshape(b) for _ in range(3)]
    if "translation" in policy:
        sx, sy = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
        out["shift"] = [torch.randint(-sx, sx + 1, size=[b, 1, 1]).numpy().reshape(b),
                        torch.randint(-sy, sy + 1, size=[b, 1, 1]).numpy().reshape(b)]
    for name, ratio in CUT.items():
        if name in policy:
            ch, cw = int(h * ratio + 0.5), int(w * ratio + 0.5)
            out["cut_hw"] = np.array([ch, cw])
            out["cut"] = [torch.randint(0, h + (1 - ch % 2), size=[b, 1, 1]).numpy().reshape(b),
                          torch.randint(0, w + (1 - cw % 2), size=[b, 1, 1]).numpy().reshape(b)]
    return out


def main():
    out = {}
    g = torch.Generator().manual_seed(51)
    x0 = torch.randn(SHAPE, generator=g)
    up = torch.randn(SHAPE, generator=g)
    for k, policy in enumerate(POLICIES):
        seed = 100 + k
        x = x0.clone().requires_grad_()
        torch.manual_seed(seed)
        y = DiffAugment(policy)(x)
        y.backward(up)
        out[f"aug{k}:y"] = y.detach().numpy()
        out[f"aug{k}:dx"] = x.grad.numpy()
        for name, vals in replay_draws(policy, SHAPE, seed).items():
            if name == "cut_hw":
                out[f"aug{k}:cut_hw"] = vals
            else:
                for i, v in enumerate(vals):
                    out[f"aug{k}:{name}{i}"] = v
    # hinge losses: two scales, values on both sides of the hinges, one exactly on a hinge
    real = [torch.randn(2, 1, 6, 6, generator=g).requires_grad_(), (torch.randn(2, 1, 3, 3, generator=g) * 2).requires_grad_()]
    fake = [torch.randn(2, 1, 6, 6, generator=g).requires_grad_(), (torch.randn(2, 1, 3, 3, generator=g) * 2).requires_grad_()]
    with torch.no_grad():
        real[0].view(-1)[0] = 1.0
        fake[0].view(-1)[0] = -1.0
    ld = discriminator_hinge_loss(real, fake)
    (ld * 3.0).backward()
    out["hinge:d_loss"] = np.float64(ld.item())
    for i in range(2):
        out[f"hinge:real{i}"] = real[i].detach().numpy(); out[f"hinge:fake{i}"] = fake[i].detach().numpy()
        out[f"hinge:d_dreal{i}"] = real[i].grad.numpy().copy(); out[f"hinge:d_dfake{i}"] = fake[i].grad.numpy().copy()
        fake[i].grad = None
    lg = generator_hinge_loss(fake)
    (lg * 3.0).backward()
    out["hinge:g_loss"] = np.float64(lg.item())
    for i in range(2):
        out[f"hinge:g_dfake{i}"] = fake[i].grad.numpy().copy()
    # non-finite discriminator outputs (a diverged D): torch.relu propagates NaN, so d_loss is NaN and train_step's
    # check raises (train_cutpp.py:326-329); relu's backward lets the upstream gradient through at a NaN input
    realn = [torch.randn(2, 1, 4, 4, generator=g), torch.randn(2, 1, 3, 3, generator=g)]
    faken = [torch.randn(2, 1, 4, 4, generator=g), torch.randn(2, 1, 3, 3, generator=g)]
    realn[0].view(-1)[3] = float("nan")
    faken[1].view(-1)[5] = float("nan")
    faken[0].view(-1)[1] = float("inf")
    realn = [t.requires_grad_() for t in realn]
    faken = [t.requires_grad_() for t in faken]
    ldn = discriminator_hinge_loss(realn, faken)
    (ldn * 3.0).backward()
    out["hinge_nan:d_loss"] = np.float64(ldn.item())
    for i in range(2):
        out[f"hinge_nan:real{i}"] = realn[i].detach().numpy(); out[f"hinge_nan:fake{i}"] = faken[i].detach().numpy()
        out[f"hinge_nan:d_dreal{i}"] = realn[i].grad.numpy().copy(); out[f"hinge_nan:d_dfake{i}"] = faken[i].grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "dside_reference.npz"), **out)
    print("wrote dside_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
