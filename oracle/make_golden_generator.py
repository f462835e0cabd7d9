"""Freeze feature maps of the UNMODIFIED reference generator and the reference loss / gradients on them.

Test infrastructure (see oracle/patchnce_oracle.py).  Run in the build container, where /root/reference is mounted:

    python oracle/make_golden_generator.py

writes tests/golden/generator_maps_b5.npz: a seeded, narrow (ngf = 8) ``ResNetGenerator``
(GAN_Variant1/models/generator_resnet_attn.py:190-235 ``get_feature_layers``) is run on two 64x64 inputs -- a "photo"
batch and a "fake" batch -- for the B5 layer ids [0, 4, 8, 12, 13]; the reference ``PatchNCELoss`` (patchnce_cut.py:25-110)
is then evaluated on those maps with ``torch.manual_seed(7)`` and differentiated w.r.t. the target maps taken as leaves
(the boundary of the rebuilt path: ``PatchNCELoss.forward(src_feats, tgt_feats)``).  The GPU test feeds the same maps
through the CUDA path (tests/test_parity_gpu.py::test_real_generator_maps_match_the_reference): real post-InstanceNorm /
post-ReLU / residual-sum statistics instead of randn, 16x16 maps where 256 draws hit every position several times.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from GAN_Variant1.losses.patchnce_cut import PatchNCELoss  # noqa: E402
from GAN_Variant1.models.generator_resnet_attn import ResNetGenerator  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
LAYERS = [0, 4, 8, 12, 13]


def main():
    torch.set_num_threads(8)
    torch.manual_seed(2024)
    gen = ResNetGenerator(ngf=8).eval()
    photo = torch.rand(2, 3, 64, 64) * 2 - 1
    fake = torch.tanh(torch.randn(2, 3, 64, 64))
    with torch.no_grad():
        src = [f.clone() for f in gen.get_feature_layers(photo, LAYERS)]
        tgt = [f.clone() for f in gen.get_feature_layers(fake, LAYERS)]
    leaves = [t.clone().requires_grad_() for t in tgt]
    crit = PatchNCELoss(0.07, 256, LAYERS)
    torch.manual_seed(7)
    loss = crit(src, leaves)
    (loss * 3.0).backward()                       # a non-unit upstream gradient
    torch.manual_seed(7)
    ids = [torch.randint(0, t.shape[2] * t.shape[3], (min(256, t.shape[2] * t.shape[3]),)) for t in tgt]   # :60-63
    d = {"loss": np.float64(loss.item()), "upstream": np.float64(3.0), "n_layers": np.int64(len(LAYERS))}
    for l, (s, t, g, i) in enumerate(zip(src, tgt, leaves, ids)):
        d[f"src{l}"], d[f"tgt{l}"] = s.numpy(), t.numpy()
        d[f"grad{l}"], d[f"ids{l}"] = g.grad.numpy(), i.numpy()
        print(l, tuple(s.shape), "grad max", float(g.grad.abs().max()), "unique ids", len(np.unique(i.numpy())))
    np.savez_compressed(os.path.join(OUT, "generator_maps_b5.npz"), **d)
    print("loss", d["loss"], os.path.getsize(os.path.join(OUT, "generator_maps_b5.npz")), "bytes")


if __name__ == "__main__":
    main()
