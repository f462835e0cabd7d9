"""D-side elementwise work in one or two launches: drop-ins for the reference's
``DiffAugment`` (``GAN_Variant1/training/diffaugment.py:62-105``) and hinge losses
(``GAN_Variant1/losses/adv_hinge.py:6-62``).

``DiffAugment(policy)(x)``: the random parameters are drawn HERE with the very calls the reference makes, in the
same order (``torch.rand(B,1,1,1, dtype=x.dtype)`` x3 for 'color', two ``torch.randint`` for 'translation', two for
'cutout'): bit-identical parameters, RNG stream aligned with everything else that draws from the device generator
(the PatchNCE ids).  The arithmetic -- brightness, saturation, contrast, the zero-padded shift, the box mask -- runs as
one pass over the image (``pnce_diffaug``; plus a per-image reduction when 'color' is on, because contrast needs the
image mean), forward and backward (the op is affine in the image: the backward needs the parameters only).  The
reference builds three (B,H,W) int64 grids, a padded NHWC copy and an advanced-indexing gather per call.

Supported policies: 'color', 'translation', 'cutout' / 'cutout_light', each at most once and in that order (both
policies the reference ships are); anything else raises.  CUDA tensors only.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib

_DT = {torch.float32: _lib.PNCE_F32, torch.float16: _lib.PNCE_F16, torch.bfloat16: _lib.PNCE_BF16}
_CUT_RATIO = {"cutout": 0.5, "cutout_light": 0.2}          # diffaugment.py:43, :60
_ORDER = {"color": 0, "translation": 1, "cutout": 2, "cutout_light": 2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _DiffAugFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, color, shift, cut, cut_hw):
        ctx.args = (color, shift, cut, cut_hw)
        return _launch(x, color, shift, cut, cut_hw, backward=False)

    @staticmethod
    def backward(ctx, g):
        color, shift, cut, cut_hw = ctx.args
        return _launch(g.contiguous(), color, shift, cut, cut_hw, backward=True), None, None, None, None


def _launch(x, color, shift, cut, cut_hw, backward):
    lib = _lib.load()
    b, c, h, w = x.shape
    y = torch.empty_like(x)
    scratch = None
    if color is not None:
        scratch = torch.empty(lib.pnce_diffaug_scratch_floats(b, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.pnce_diffaug(
        x.data_ptr(), y.data_ptr(), _DT[x.dtype], b, c, h, w,
        _ptr(color[0]) if color else None, _ptr(color[1]) if color else None, _ptr(color[2]) if color else None,
        _ptr(shift[0]) if shift else None, _ptr(shift[1]) if shift else None,
        _ptr(cut[0]) if cut else None, _ptr(cut[1]) if cut else None, cut_hw[0], cut_hw[1],
        _ptr(scratch), 1 if backward else 0, torch.cuda.current_stream(x.device).cuda_stream), "pnce_diffaug")
    return y


class DiffAugment:
    """Drop-in for ``DiffAugment`` -- diffaugment.py:62-105 (same constructor, same ``__call__``)."""

    def __init__(self, policy: Optional[Sequence[str]] = None):
        if policy is None:
            policy = ["color", "translation", "cutout_light"]                 # :72-73
        self.policy = list(policy)
        known = [p for p in self.policy if p in _ORDER]                        # unknown names are skipped, :77-79
        ranks = [_ORDER[p] for p in known]
        if ranks != sorted(set(ranks)):
            raise NotImplementedError(f"DiffAugment policy {self.policy}: the fused kernel applies color -> translation "
                                      "-> cutout, each at most once (use the reference's DiffAugment for other orders)")
        self._color = "color" in known
        self._translation = "translation" in known
        self._cut_ratio = next((_CUT_RATIO[p] for p in known if p in _CUT_RATIO), None)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not (self._color or self._translation or self._cut_ratio is not None):
            return x
        if not x.is_cuda:
            raise RuntimeError("DiffAugment: CUDA tensors only (the B200 path has no CPU fallback)")
        if x.dim() != 4 or x.dtype not in _DT:
            raise RuntimeError(f"DiffAugment: expected a (B, C, H, W) float tensor, got {tuple(x.shape)} {x.dtype}")
        b, _, h, w = x.shape
        dev = x.device
        color = shift = cut = None
        cut_hw = (0, 0)
        if self._color:                                                        # :8, :15, :22 -- one draw each, in order
            color = tuple(torch.rand(b, 1, 1, 1, dtype=x.dtype, device=dev) for _ in range(3))
        if self._translation:                                                  # :27-29
            sx, sy = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
            shift = (torch.randint(-sx, sx + 1, size=[b, 1, 1], device=dev),
                     torch.randint(-sy, sy + 1, size=[b, 1, 1], device=dev))
        if self._cut_ratio is not None:                                        # :45-47
            cut_hw = (int(h * self._cut_ratio + 0.5), int(w * self._cut_ratio + 0.5))
            if cut_hw[0] <= 0 or cut_hw[1] <= 0:
                raise RuntimeError("DiffAugment: image too small for the cutout ratio")
            cut = (torch.randint(0, h + (1 - cut_hw[0] % 2), size=[b, 1, 1], device=dev),
                   torch.randint(0, w + (1 - cut_hw[1] % 2), size=[b, 1, 1], device=dev))
        return _DiffAugFn.apply(x.contiguous(), color, shift, cut, cut_hw)


# ------------------------------------------------------------------------------------------------------------------
class _HingeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, n_scales, *preds):
        lib = _lib.load()
        preds = [p.contiguous() for p in preds]
        real = preds[:n_scales] if mode == 0 else None
        fake = preds[n_scales:] if mode == 0 else preds
        t0 = fake[0]
        dev = t0.device
        for p in preds:
            if not p.is_cuda or p.dtype != t0.dtype or p.device != dev:
                raise RuntimeError("hinge loss: predictions must be CUDA tensors of one dtype on one device "
                                   "(the B200 path has no CPU fallback)")
        if mode == 0:
            for r, f in zip(real, fake):
                if r.numel() != f.numel():
                    raise RuntimeError("hinge loss: real / fake predictions of a scale differ in size")
        vp = ctypes.c_void_p * n_scales
        numel = (ctypes.c_longlong * n_scales)(*[f.numel() for f in fake])
        rp = vp(*[r.data_ptr() for r in real]) if mode == 0 else None
        fp = vp(*[f.data_ptr() for f in fake])
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(lib.pnce_hinge_fwd(rp, fp, numel, n_scales, mode, _DT[t0.dtype], loss.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), "pnce_hinge_fwd")
        ctx.save_for_backward(*preds)
        ctx.mode, ctx.n_scales = mode, n_scales
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        preds = ctx.saved_tensors
        mode, n = ctx.mode, ctx.n_scales
        real = preds[:n] if mode == 0 else None
        fake = preds[n:] if mode == 0 else preds
        dev = fake[0].device
        need = ctx.needs_input_grad[2:]
        grads: List[Optional[torch.Tensor]] = [torch.empty_like(p) if nd else None for p, nd in zip(preds, need)]
        vp = ctypes.c_void_p * n
        numel = (ctypes.c_longlong * n)(*[f.numel() for f in fake])
        rp = vp(*[r.data_ptr() for r in real]) if mode == 0 else None
        fp = vp(*[f.data_ptr() for f in fake])
        dr = vp(*[_ptr(x) for x in grads[:n]]) if mode == 0 else None
        df = vp(*[_ptr(x) for x in (grads[n:] if mode == 0 else grads)])
        g32 = g.detach().to(device=dev, dtype=torch.float32).contiguous()
        _lib.check(lib.pnce_hinge_bwd(rp, fp, dr, df, numel, n, mode, _DT[fake[0].dtype], g32.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), "pnce_hinge_bwd")
        return (None, None, *grads)


def discriminator_hinge_loss(real_preds, fake_preds):
    """Drop-in for adv_hinge.py:6-32: mean over scales of 0.5 * (mean relu(1 - real) + mean relu(1 + fake))."""
    if not isinstance(real_preds, list):                                       # :19-21
        real_preds, fake_preds = [real_preds], [fake_preds]
    n = min(len(real_preds), len(fake_preds))                                  # zip() truncates, :23
    if n == 0:
        raise ZeroDivisionError("float division by zero")                      # loss / len(real_preds), :31
    loss = _HingeFn.apply(0, n, *real_preds[:n], *fake_preds[:n])
    return loss if n == len(real_preds) else loss * (n / len(real_preds))


def generator_hinge_loss(fake_preds):
    """Drop-in for adv_hinge.py:35-62: mean over scales of -mean(fake)."""
    if not isinstance(fake_preds, list):
        fake_preds = [fake_preds]
    if len(fake_preds) == 0:
        raise ZeroDivisionError("float division by zero")
    return _HingeFn.apply(1, len(fake_preds), *fake_preds)
