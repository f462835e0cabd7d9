"""CPU oracle for the D-side row.  TEST INFRASTRUCTURE ONLY (same rules as
patchnce_oracle.py).  Restates in float64 numpy, with the random parameters given:

* DiffAugment, training/diffaugment.py: rand_brightness :6-9, rand_saturation :12-16, rand_contrast :19-23,
  rand_translation :26-37 (zero-padded shift), rand_cutout :40-57 (box mask with clamped indices), applied in policy
  order color -> translation -> cutout (:99-105), and the vector-Jacobian product of the whole chain;
* the hinge losses, losses/adv_hinge.py:6-62, and their gradients.

Pinned by tests/golden/dside_reference.npz, frozen from the unmodified reference by oracle/make_golden_dside.py.
"""
import numpy as np


def _cut_mask(b, h, w, ox, oy, cut_hw):
    mask = np.ones((b, h, w))
    ch, cw = cut_hw
    for n in range(b):
        gx = np.clip(np.arange(ch) + int(ox[n]) - ch // 2, 0, h - 1)           # :52
        gy = np.clip(np.arange(cw) + int(oy[n]) - cw // 2, 0, w - 1)           # :53
        mask[np.ix_([n], gx, gy)] = 0                                          # :55
    return mask


def _shift(x, tx, ty, inverse=False):
    """y[b,:,i,j] = x[b,:,i+tx,j+ty] or 0 outside (:30-36); inverse=True is the adjoint (scatter back)."""
    b, c, h, w = x.shape
    y = np.zeros_like(x)
    for n in range(b):
        dx, dy = int(tx[n]), int(ty[n])
        for i in range(h):
            u = i + dx
            if not 0 <= u < h:
                continue
            for j in range(w):
                v = j + dy
                if 0 <= v < w:
                    if inverse:
                        y[n, :, u, v] = x[n, :, i, j]
                    else:
                        y[n, :, i, j] = x[n, :, u, v]
    return y


def diffaug_np(x, color=None, shift=None, cut=None, cut_hw=(0, 0)):
    """color = (rb, rs, rc) each (B,), shift = (tx, ty), cut = (ox, oy); returns the augmented images (float64)."""
    x = np.asarray(x, np.float64)
    b, c, h, w = x.shape
    if color is not None:
        rb, rs, rc = (np.asarray(t, np.float64).reshape(b, 1, 1, 1) for t in color)
        x = x + (rb - 0.5)                                                     # :8
        m = x.mean(axis=1, keepdims=True)
        x = (x - m) * (rs * 2) + m                                             # :14-15
        mu = x.mean(axis=(1, 2, 3), keepdims=True)
        x = (x - mu) * (rc + 0.5) + mu                                         # :21-22
    if shift is not None:
        x = _shift(x, np.asarray(shift[0]).reshape(-1), np.asarray(shift[1]).reshape(-1))
    if cut is not None:
        x = x * _cut_mask(b, h, w, np.asarray(cut[0]).reshape(-1), np.asarray(cut[1]).reshape(-1), cut_hw)[:, None]
    return x


def diffaug_vjp_np(g, color=None, shift=None, cut=None, cut_hw=(0, 0)):
    """d loss / d input given g = d loss / d output (the chain is affine in the image)."""
    g = np.asarray(g, np.float64)
    b, c, h, w = g.shape
    if cut is not None:
        g = g * _cut_mask(b, h, w, np.asarray(cut[0]).reshape(-1), np.asarray(cut[1]).reshape(-1), cut_hw)[:, None]
    if shift is not None:
        g = _shift(g, np.asarray(shift[0]).reshape(-1), np.asarray(shift[1]).reshape(-1), inverse=True)
    if color is not None:
        _, rs, rc = (np.asarray(t, np.float64).reshape(b, 1, 1, 1) for t in color)
        sc, ss = rc + 0.5, rs * 2
        g = sc * g + (1 - sc) * g.mean(axis=(1, 2, 3), keepdims=True)
        g = ss * g + (1 - ss) * g.mean(axis=1, keepdims=True)
    return g


def d_hinge_np(real, fake):
    """(loss, [d real], [d fake]) of discriminator_hinge_loss, adv_hinge.py:6-32."""
    s = len(real)
    loss = sum(0.5 * (np.maximum(1 - r, 0).mean() + np.maximum(1 + f, 0).mean()) for r, f in zip(real, fake)) / s
    # torch.relu: NaN propagates in the forward (np.maximum does the same) and its backward passes the gradient at NaN
    dr = [np.where(~(r >= 1), -0.5 / (s * r.size), 0.0) for r in real]
    df = [np.where(~(f <= -1), 0.5 / (s * f.size), 0.0) for f in fake]
    return float(loss), dr, df


def g_hinge_np(fake):
    """(loss, [d fake]) of generator_hinge_loss, adv_hinge.py:35-62."""
    s = len(fake)
    return float(sum(-f.mean() for f in fake) / s), [np.full(f.shape, -1.0 / (s * f.size)) for f in fake]
