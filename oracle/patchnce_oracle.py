"""CPU oracle for the PatchNCE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module, and only as the checker (or as the timed CPU baseline).  The
product package ``gan_variant_research_b200`` never imports it and has no CPU fallback.

What it restates (all citations relative to the upstream reference tree,
``GAN_Variant1/losses/patchnce_cut.py`` unless another file is named):

* the patch-id draw                                   :60-63
* the NCHW -> (B, HW, C) view + per-image gather      :53-74
* L2 normalisation ``x / max(||x||, 1e-6)``           :77-78
* per-image logits ``q k^T / tau`` clamped to +-50    :83-88
* diagonal-positive cross entropy (mean over rows)    :91-94
* the non-finite guards (per image, per layer)        :97-110
* the layer mean ``sum_l / len(src_feats)``           :34-40
* the wrapper that extracts src (no grad) / tgt feats :113-149

Parity pin: ``oracle/make_golden.py`` imports the *unmodified* reference from ``/root/reference`` in
the build container, runs it on seeded inputs and freezes ids / losses / gradients under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against those files.
The reference itself has no tests or golden vectors, so "reference run here" is the pin.

The netF head (``head_*`` functions at the bottom) has no counterpart in the reference:
PARITY UNPINNED for that sub-piece -- it is a plain restatement of Linear -> ReLU -> Linear
(north_star piece 3) and is only self-consistent.

Two independent restatements are kept on purpose:

* ``*_torch``  -- op-for-op with autograd (this is also what the CPU baseline times: it issues
  the same ATen op sequence as the reference, including the per-image Python loops whose
  backward is O(B^2) dense traffic);
* ``*_np``     -- float64 numpy with the backward written out analytically, which is the formula
  sheet the CUDA kernels implement.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NORM_EPS = 1e-6      # F.normalize(..., eps=1e-6)                       :77-78
LOGIT_CLAMP = 50.0   # torch.clamp(logits, -50, 50)                     :88


# --------------------------------------------------------------------------------------------
# id draw
# --------------------------------------------------------------------------------------------
def patch_count(num_patches: int, hw: int) -> int:
    """P = min(num_patches, H*W)                                          :60"""
    return min(int(num_patches), int(hw))


def draw_patch_ids(hw: int, num_patches: int, device="cpu") -> torch.Tensor:
    """One draw per layer, with replacement, shared by every image, on the global generator of
    the features' device.                                                  :60-63"""
    return torch.randint(0, hw, (patch_count(num_patches, hw),), device=device)


def mt19937_ids(seed: int, hws, num_patches: int):
    """CPU id law (verified against the reference): ids are consecutive raw
    32-bit outputs of mt19937(seed) reduced ``% HW``, consumed layer after layer."""
    rs = np.random.RandomState()
    # torch seeds its CPU mt19937 with init_genrand(seed); numpy's legacy seeding of a 32-bit
    # integer is the same init_genrand.
    rs.seed(np.uint32(seed))
    out = []
    for hw in hws:
        p = patch_count(num_patches, hw)
        raw = rs.randint(0, 2**32, size=p, dtype=np.uint64)
        out.append((raw % np.uint64(hw)).astype(np.int64))
    return out


# --------------------------------------------------------------------------------------------
# torch restatement (autograd supplies the backward)
# --------------------------------------------------------------------------------------------
def layer_loss_torch(src: torch.Tensor, tgt: torch.Tensor, ids: torch.Tensor,
                     temperature: float, warn=None) -> torch.Tensor:
    """One layer of PatchNCE for given ids.                               :42-110

    ``src``/``tgt``: (B, C, H, W).  Returns the layer loss (mean over the batch of the per-image
    diagonal CE).  ``warn`` collects the messages the reference would print."""
    b_sz, c_sz = src.shape[0], src.shape[1]
    k_map = src.reshape(b_sz, c_sz, -1).transpose(1, 2)          # (B, HW, C) view   :56
    q_map = tgt.reshape(b_sz, c_sz, -1).transpose(1, 2)          #                   :57
    k_rows = torch.stack([k_map[b][ids] for b in range(b_sz)])   # (B, P, C)         :69-74
    q_rows = torch.stack([q_map[b][ids] for b in range(b_sz)])
    k_rows = F.normalize(k_rows, dim=2, eps=NORM_EPS)            #                   :77
    q_rows = F.normalize(q_rows, dim=2, eps=NORM_EPS)            #                   :78
    n_rows = ids.numel()
    acc = 0.0
    for b in range(b_sz):                                        #                   :83
        z = (q_rows[b] @ k_rows[b].t()) / temperature            #                   :85
        z = z.clamp(-LOGIT_CLAMP, LOGIT_CLAMP)                   #                   :88
        target = torch.arange(n_rows, device=z.device)           #                   :91
        l_b = F.cross_entropy(z, target, reduction="mean")       #                   :94
        if not torch.isfinite(l_b):                              #                   :97-99
            if warn is not None:
                warn.append(("image", b))
            l_b = torch.zeros((), device=z.device)
        acc = acc + l_b                                          #                   :101
    out = acc / b_sz                                             #                   :103
    if not torch.isfinite(out):                                  #                   :106-108
        if warn is not None:
            warn.append(("layer", -1))
        return torch.zeros((), device=out.device, requires_grad=True)
    return out


def patchnce_loss_torch(src_feats, tgt_feats, temperature=0.07, num_patches=256,
                        ids_list=None, warn=None):
    """``PatchNCELoss(temperature, num_patches, nce_layers).forward``      :25-40

    When ``ids_list`` is None the ids are drawn exactly like the reference (one ``randint`` per
    zipped layer, in layer order).  Returns ``(loss, ids_list)``."""
    used = []
    total = 0.0
    for li, (s, t) in enumerate(zip(src_feats, tgt_feats)):
        hw = s.shape[2] * s.shape[3]
        ids = draw_patch_ids(hw, num_patches, s.device) if ids_list is None else ids_list[li]
        used.append(ids)
        total = total + layer_loss_torch(s, t, ids, temperature, warn)
    return total / len(src_feats), used                          #                   :40


def compute_patchnce_loss_torch(generator, src_images, tgt_images, nce_layers,
                                temperature=0.07, num_patches=256, ids_list=None):
    """``compute_patchnce_loss``                                           :113-149
    (``generator.get_feature_layers`` is models/generator_resnet_attn.py:190-235)."""
    with torch.no_grad():                                        #                   :138-139
        src_feats = generator.get_feature_layers(src_images, nce_layers)
    src_feats = [f.detach() for f in src_feats]                  #                   :142
    tgt_feats = generator.get_feature_layers(tgt_images, nce_layers)   #             :145
    loss, _ = patchnce_loss_torch(src_feats, tgt_feats, temperature, num_patches, ids_list)
    return loss


# --------------------------------------------------------------------------------------------
# numpy float64 restatement with the analytic backward
# --------------------------------------------------------------------------------------------
def _normalize_np(x):
    """x / max(||x||, eps) along the last axis; also returns the raw norms."""
    n = np.sqrt((x * x).sum(-1, keepdims=True))
    return x / np.maximum(n, NORM_EPS), n


def layer_loss_and_grad_np(src, tgt, ids, temperature, n_layers=1, upstream=1.0):
    """float64 forward + analytic backward of one layer.

    Returns ``(layer_loss, per_image_losses, dense_grad_tgt)`` where ``dense_grad_tgt`` is
    d(total)/d(tgt) with total = upstream * (sum_l layer_loss_l) / n_layers, i.e. the factor
    1/(P * B * L) of SURVEY.md section 3.2 is applied here.  Non-finite per-image losses contribute 0
    and a zero upstream gradient (:97-99) -- see the NaN note in the loop."""
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    ids = np.asarray(ids, dtype=np.int64)
    b_sz, c_sz, h, w = src.shape
    p = ids.shape[0]
    k_raw = src.reshape(b_sz, c_sz, h * w)[:, :, ids].transpose(0, 2, 1)   # (B,P,C)
    q_raw = tgt.reshape(b_sz, c_sz, h * w)[:, :, ids].transpose(0, 2, 1)
    with np.errstate(all="ignore"):
        k_hat, _ = _normalize_np(k_raw)
        q_hat, q_n = _normalize_np(q_raw)
        grad = np.zeros((b_sz, c_sz, h * w), dtype=np.float64)
        per_image = np.zeros(b_sz, dtype=np.float64)
        for b in range(b_sz):
            z_raw = (q_hat[b] @ k_hat[b].T) / temperature
            z = np.clip(z_raw, -LOGIT_CLAMP, LOGIT_CLAMP)
            m = z.max(axis=1, keepdims=True)
            e = np.exp(z - m)
            s = e.sum(axis=1, keepdims=True)
            row_loss = (np.log(s) + m)[:, 0] - np.diag(z)
            l_b = row_loss.mean()
            if not np.isfinite(l_b):
                # 0 loss (:97-99).  The loss node is replaced by a constant, so the image's rows
                # receive an exactly-zero upstream gradient -- but autograd still runs the
                # F.normalize backward on them (0 / NaN, NaN * 0), so every channel of a sampled
                # target patch that holds a NaN/Inf comes out NaN; all other entries are 0.
                # (Observed in the reference run frozen as tests/golden/small_nan_image.npz.)
                bad_rows = ~np.isfinite(q_raw[b]).all(axis=1)
                if bad_rows.any():
                    grad[b][:, ids[bad_rows]] = np.nan
                continue
            per_image[b] = l_b
            # torch.clamp backward passes the gradient where min <= x <= max (inclusive).
            mask = (z_raw >= -LOGIT_CLAMP) & (z_raw <= LOGIT_CLAMP)
            d_z = (e / s - np.eye(p)) * mask * (upstream / (p * b_sz * n_layers))
            d_qhat = (d_z @ k_hat[b]) / temperature                # dK never needed: src detached
            # F.normalize backward: n = max(||x||, eps); the norm term only flows when ||x|| >= eps
            n_b = q_n[b]
            big = n_b >= NORM_EPS
            proj = (q_hat[b] * d_qhat).sum(-1, keepdims=True)
            d_x = np.where(big, (d_qhat - q_hat[b] * proj) / np.maximum(n_b, NORM_EPS),
                           d_qhat / NORM_EPS)
            np.add.at(grad[b].T, ids, d_x)                          # duplicate ids accumulate
    layer_loss = per_image.sum() / b_sz
    return layer_loss, per_image, grad.reshape(b_sz, c_sz, h, w)


def patchnce_loss_and_grads_np(src_feats, tgt_feats, ids_list, temperature=0.07, upstream=1.0):
    """All layers: ``(loss, [layer losses], [dense grads])`` in float64."""
    n_layers = len(src_feats)
    losses, grads = [], []
    for s, t, ids in zip(src_feats, tgt_feats, ids_list):
        l, _, g = layer_loss_and_grad_np(s, t, ids, temperature, n_layers, upstream)
        losses.append(l)
        grads.append(g)
    return float(np.sum(losses) / n_layers), losses, grads


def gather_normalize_np(feat, ids):
    """(B,C,H,W) -> normalised rows (B*P, C) float64 and raw norms (B*P,): the PatchSampleF
    (use_mlp=False) contract of SURVEY.md section 8b, restating :53-78."""
    feat = np.asarray(feat, dtype=np.float64)
    b_sz, c_sz, h, w = feat.shape
    rows = feat.reshape(b_sz, c_sz, h * w)[:, :, np.asarray(ids)].transpose(0, 2, 1)
    hat, n = _normalize_np(rows)
    return hat.reshape(-1, c_sz), n.reshape(-1)


# --------------------------------------------------------------------------------------------
# netF head -- PARITY UNPINNED by the reference (it has no MLP); north_star piece (3)
# --------------------------------------------------------------------------------------------
def head_forward_torch(rows, w1, b1, w2, b2):
    """Linear(C,nc) -> ReLU -> Linear(nc,nc) -> x / max(||x||, 1e-6) on (N, C) rows."""
    h = torch.relu(rows @ w1.t() + b1)
    y = h @ w2.t() + b2
    return F.normalize(y, dim=1, eps=NORM_EPS)


def patchnce_head_loss_torch(src_feats, tgt_feats, ids_list, heads, temperature=0.07):
    """PatchNCE with a per-layer head applied to the *raw* gathered rows of both q and k
    (upstream-CUT ordering: gather -> MLP -> L2 norm), then the reference's logits / diagonal CE
    (:83-103).  ``heads`` is a list of (w1, b1, w2, b2).  src features carry no gradient."""
    total = 0.0
    for s, t, ids, (w1, b1, w2, b2) in zip(src_feats, tgt_feats, ids_list, heads):
        b_sz, c_sz = s.shape[0], s.shape[1]
        k_rows = s.reshape(b_sz, c_sz, -1).transpose(1, 2)[:, ids, :]
        q_rows = t.reshape(b_sz, c_sz, -1).transpose(1, 2)[:, ids, :]
        p = ids.numel()
        # upstream CUT detaches feat_k inside the loss: the head only learns through q
        k_hat = head_forward_torch(k_rows.reshape(-1, c_sz), w1, b1, w2, b2).reshape(b_sz, p, -1).detach()
        q_hat = head_forward_torch(q_rows.reshape(-1, c_sz), w1, b1, w2, b2).reshape(b_sz, p, -1)
        acc = 0.0
        for b in range(b_sz):
            z = (q_hat[b] @ k_hat[b].t() / temperature).clamp(-LOGIT_CLAMP, LOGIT_CLAMP)
            acc = acc + F.cross_entropy(z, torch.arange(p, device=z.device))
        total = total + acc / b_sz
    return total / len(src_feats)
