"""CPU oracle for the AMP optimiser step.  TEST INFRASTRUCTURE ONLY (same rules as
patchnce_oracle.py).  Restates, in float32 numpy, what the reference's AMPContext.step_optimizer does to an
optim.Adam -- utils/amp_utils.py:29-41:

* scaler.unscale_(optimizer): found_inf = any non-finite RAW gradient; grad *= fl32(1 / scale)             :33
* clip_grad_norm_(params with a gradient, max_norm): total = || per-tensor norms ||_2;
  coef = min(max_norm / (total + 1e-6), 1); grad *= coef                                                   :35-38
* scaler.step(optimizer): optimizer.step() unless found_inf                                                :40
  Adam (torch.optim.adam, default single-group path): step += 1; m = lerp(m, g, 1 - b1);
  v = v * b2 + (1 - b2) * g * g; p += -(lr / (1 - b1^step)) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
* scaler.update(): found_inf -> scale *= backoff, tracker = 0; else tracker += 1 and, on reaching the growth
  interval, scale *= growth, tracker = 0                                                                   :41

Pinned by tests/golden/amp_step_reference.npz, frozen from the unmodified reference by oracle/make_golden_amp_step.py.
"""
import math

import numpy as np

F = np.float32


def amp_adam_step_np(params, grads, exp_avg, exp_avg_sq, steps, scale, tracker, *, lr, betas, eps=1e-8,
                     weight_decay=0.0, max_grad_norm=None, growth_factor=2.0, backoff_factor=0.5,
                     growth_interval=2000):
    """One step_optimizer call on lists of float32 arrays (updated in place).  ``scale`` None = no scaler.
    Returns (scale, tracker, total_norm)."""
    found_inf = any(not np.isfinite(g).all() for g in grads)
    with np.errstate(all="ignore"):
        if scale is not None:
            inv = F(1.0 / float(scale))
            for g in grads:
                g *= inv
        total = None
        if max_grad_norm is not None:
            norms = np.array([np.sqrt(np.sum(g.astype(np.float64) ** 2)) for g in grads]).astype(F)
            total = F(np.sqrt(np.sum(norms.astype(np.float64) ** 2)))
            coef = F(max_grad_norm) / (total + F(1e-6))
            coef = coef if np.isnan(coef) else min(coef, F(1.0))
            for g in grads:
                g *= F(coef)
        if not (found_inf and scale is not None):
            b1, b2 = betas
            for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avg, exp_avg_sq)):
                steps[i] += 1
                gg = g if weight_decay == 0 else (g + F(weight_decay) * p).astype(F)
                w = F(1.0 - b1)
                m[...] = (m + w * (gg - m)).astype(F) if abs(w) < 0.5 else (gg - (gg - m) * (F(1) - w)).astype(F)
                v[...] = ((v * F(b2)).astype(F) + F(1.0 - b2) * (gg * gg).astype(F)).astype(F)
                bc1 = 1.0 - b1 ** steps[i]
                bc2s = math.sqrt(1.0 - b2 ** steps[i])
                den = ((np.sqrt(v) / F(bc2s)).astype(F) + F(eps)).astype(F)
                p += (F(-(lr / bc1)) * (m / den).astype(F)).astype(F)
    if scale is not None:
        if found_inf:
            scale, tracker = F(scale * F(backoff_factor)), 0
        else:
            tracker += 1
            if tracker == growth_interval:
                ns = F(scale * F(growth_factor))
                scale = ns if np.isfinite(ns) else scale
                tracker = 0
    return scale, tracker, total
