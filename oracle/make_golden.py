"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The fixtures are what pins the oracle (and through it the CUDA path) to the reference; the
reference has no golden vectors of its own.  Nothing here is imported at test time.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PNCE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from GAN_Variant1.losses.patchnce_cut import PatchNCELoss, compute_patchnce_loss  # noqa: E402
from GAN_Variant1.models.generator_resnet_attn import ResNetGenerator  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
R4 = [(64, 256, 256), (256, 64, 64), (256, 64, 64), (128, 128, 128)]


def capture_ids(fn):
    """Run fn() while recording every torch.randint result (the reference's only RNG use)."""
    rec = []
    orig = torch.randint

    def spy(*a, **k):
        r = orig(*a, **k)
        rec.append(r.clone())
        return r
    torch.randint = spy
    try:
        out = fn()
    finally:
        torch.randint = orig
    return out, rec


def make_feats(shapes, b, seed, relu_mask=None):
    g = torch.Generator().manual_seed(seed)
    relu_mask = relu_mask or [True] * len(shapes)
    src = [torch.randn(b, *s, generator=g) for s in shapes]
    src = [x.relu() if r else x for x, r in zip(src, relu_mask)]
    tgt = [torch.randn(b, *s, generator=g) for s in shapes]
    tgt = [x.relu() if r else x for x, r in zip(tgt, relu_mask)]
    return src, [t.requires_grad_() for t in tgt]


def survey_case(b):
    """full-size R4 maps, data seed 1234, id seed 7."""
    src, tgt = make_feats(R4, b, 1234)
    torch.manual_seed(7)
    mod = PatchNCELoss(0.07, 256, [0, 4, 8, 12, 16])
    per_layer = []
    orig = mod._compute_nce_loss

    def spy_layer(s, t):
        l = orig(s, t)
        per_layer.append(float(l))
        return l
    mod._compute_nce_loss = spy_layer
    loss, ids = capture_ids(lambda: mod(src, tgt))
    loss.backward()
    out = {"loss": np.float64(loss.item()), "per_layer": np.array(per_layer)}
    for i, t in enumerate(tgt):
        g = t.grad
        out[f"ids{i}"] = ids[i].numpy()
        out[f"gnorm{i}"] = np.float64(g.double().norm().item())
        out[f"gsum{i}"] = np.float64(g.double().sum().item())
        out[f"nnz{i}"] = np.int64((g != 0).sum().item())
        # gradient columns at the first 8 sampled positions of image 0 (all channels)
        cols = ids[i][:8]
        out[f"gcols{i}"] = g[0].reshape(g.shape[1], -1)[:, cols].numpy().copy()
    return out


def small_case(name, shapes, b, num_patches, seed, id_seed, tau=0.07, relu_mask=None,
               mutate=None, upstream=1.0):
    src, tgt = make_feats(shapes, b, seed, relu_mask)
    if mutate is not None:
        with torch.no_grad():
            mutate(src, tgt)
    torch.manual_seed(id_seed)
    mod = PatchNCELoss(tau, num_patches, list(range(len(shapes))))
    loss, ids = capture_ids(lambda: mod(src, tgt))
    if loss.requires_grad:
        (loss * upstream).backward()
    out = {"loss": np.float64(loss.item()), "tau": np.float64(tau), "upstream": np.float64(upstream),
           "num_patches": np.int64(num_patches), "n_layers": np.int64(len(shapes))}
    for i, (s, t) in enumerate(zip(src, tgt)):
        out[f"src{i}"] = s.numpy()
        out[f"tgt{i}"] = t.detach().numpy()
        out[f"ids{i}"] = ids[i].numpy()
        out[f"grad{i}"] = (t.grad if t.grad is not None else torch.zeros_like(t)).numpy()
    np.savez_compressed(os.path.join(OUT, f"small_{name}.npz"), **out)
    print(name, "loss", out["loss"], [ids[i].numel() for i in range(len(shapes))])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # ---- 1. survey tripwire, B=1 and B=2 -------------------------------------------------
    for b in (1, 2):
        d = survey_case(b)
        np.savez_compressed(os.path.join(OUT, f"survey_r4_b{b}.npz"), **d)
        print("survey b", b, d["loss"], d["per_layer"], [d[f"gnorm{i}"] for i in range(4)])

    # ---- 2. small cases with full tensors ------------------------------------------------
    small_case("basic", [(8, 6, 6), (16, 4, 4)], 2, 16, seed=11, id_seed=3)
    # P clipped to HW (num_patches > HW) and ragged channel counts / odd map sizes
    small_case("ragged", [(3, 5, 7), (20, 3, 3), (33, 8, 8)], 3, 64, seed=12, id_seed=4,
               relu_mask=[True, False, True])
    # AMP-like upstream gradient (GradScaler 65536 x lambda 10)
    small_case("scaled", [(16, 8, 8)], 2, 32, seed=13, id_seed=5, upstream=65536.0 * 10.0)
    # low temperature so the +-50 clamp engages (1/tau = 100)
    small_case("clamp", [(12, 6, 6)], 2, 24, seed=14, id_seed=6, tau=0.01, relu_mask=[False])

    # all-zero target/source patches (post-ReLU dead vectors): dx = g / eps branch
    def zero_patch(src, tgt):
        torch.manual_seed(8)
        ids = torch.randint(0, 36, (16,))
        tgt[0][0].reshape(8, -1)[:, ids[2]] = 0.0
        src[0][1].reshape(8, -1)[:, ids[5]] = 0.0
        tgt[0][1].reshape(8, -1)[:, ids[5]] = 0.0
    small_case("zerovec", [(8, 6, 6)], 2, 16, seed=15, id_seed=8, mutate=zero_patch)

    # NaN in one image at a sampled position: that image contributes 0 loss / 0 grad (:97-99)
    def nan_patch(src, tgt):
        torch.manual_seed(9)
        ids = torch.randint(0, 36, (16,))
        tgt[0][1].reshape(8, -1)[3, ids[0]] = float("nan")
    small_case("nan_image", [(8, 6, 6), (4, 6, 6)], 3, 16, seed=16, id_seed=9, mutate=nan_patch)

    # Inf in the source of one image
    def inf_patch(src, tgt):
        torch.manual_seed(10)
        ids = torch.randint(0, 16, (16,))
        src[0][0].reshape(8, -1)[1, ids[4]] = float("inf")
    small_case("inf_src", [(8, 4, 4)], 2, 16, seed=17, id_seed=10, mutate=inf_patch)

    # medium case exercising P=256 with duplicates on a 64x64 map, C=64, B=2 (full tensors: ~4 MB raw)
    small_case("p256", [(64, 32, 32)], 2, 256, seed=18, id_seed=11)

    # ---- 3. end-to-end through the reference generator ----------------------------------
    e2e = {}
    for tag, layers in (("r4", [0, 4, 8, 12, 16]), ("b5", [0, 4, 8, 12, 13])):
        torch.manual_seed(0)
        gen = ResNetGenerator()
        x = torch.randn(1, 3, 256, 256)
        y = torch.tanh(torch.randn(1, 3, 256, 256)).requires_grad_()
        torch.manual_seed(7)
        loss = compute_patchnce_loss(gen, x, y, layers, 0.07, 256)
        loss.backward()
        e2e[f"loss_{tag}"] = np.float64(loss.item())
        e2e[f"gnorm_{tag}"] = np.float64(y.grad.double().norm().item())
        print("e2e", tag, e2e[f"loss_{tag}"], e2e[f"gnorm_{tag}"])
    np.savez_compressed(os.path.join(OUT, "e2e_generator.npz"), **e2e)

    # ---- 4. CPU id law: ids for seeds / sizes ---------------------------------------------
    law = {}
    for seed in (0, 7, 12345):
        torch.manual_seed(seed)
        for j, hw in enumerate((65536, 4096, 4096, 16384, 100)):
            law[f"s{seed}_l{j}"] = torch.randint(0, hw, (min(256, hw),)).numpy()
    np.savez_compressed(os.path.join(OUT, "cpu_id_law.npz"), **law)


if __name__ == "__main__":
    main()
