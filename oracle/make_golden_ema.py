"""Golden fixture for the EMA row from the UNMODIFIED reference:
utils/io_ckpt.py::EMA driven through a few parameter updates.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PNCE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "GAN_Variant1"))
sys.path.insert(0, REF)
from GAN_Variant1.utils.io_ckpt import EMA  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def make_model():
    torch.manual_seed(31)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.InstanceNorm2d(7, affine=True),
                              torch.nn.Conv2d(7, 5, 1, bias=False), torch.nn.Linear(9, 4))
    net[2].weight.requires_grad_(False)          # frozen parameters are skipped by the reference (:20, :26)
    return net


def main():
    net = make_model()
    ema = EMA(net, decay=0.999)
    out = {"decay": np.float64(0.999)}
    g = torch.Generator().manual_seed(32)
    for step in range(4):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        ema.update()
        for name, v in ema.shadow.items():
            out[f"s{step}:{name}"] = v.numpy().copy()
    for name, p in net.named_parameters():
        out[f"final:{name}"] = p.detach().numpy().copy()
    ema.apply_shadow()
    for name, p in net.named_parameters():
        out[f"applied:{name}"] = p.detach().numpy().copy()
    ema.restore()
    for name, p in net.named_parameters():
        assert np.array_equal(p.detach().numpy(), out[f"final:{name}"])
    np.savez_compressed(os.path.join(OUT, "ema_reference.npz"), **out)
    print("wrote ema_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
