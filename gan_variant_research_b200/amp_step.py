"""The AMP optimiser step in three launches: drop-in for the reference's
``AMPContext.step_optimizer(optimizer, max_grad_norm)`` -- ``GAN_Variant1/utils/amp_utils.py:29-41`` -- on the
``optim.Adam`` that ``training/sched_optim.py:20-25`` builds.

What the reference runs per call, for the generator's 48 parameter tensors: ``scaler.unscale_`` (a multi-tensor
launch per 3 dtypes/devices + a found-inf reduction), ``clip_grad_norm_`` (per-tensor norms, a stack, a norm, a clamp,
a multi-tensor multiply), ``scaler.step`` (**a device-to-host sync**, ``found_inf.item()``, to decide whether to skip;
then foreach-Adam: seven multi-tensor ops, each split into several launches) and ``scaler.update``.  Here:
``k_amp_gradnorm`` (non-finite check + partial sums of squares), ``k_amp_adam`` (unscale, clip, Adam) and
``k_amp_finish`` (step counters, loss-scale update) through ``pnce_amp_adam_step`` (include/pnce.h) -- no host sync,
since the skip decision never leaves the device.  The arithmetic follows ATen's foreach Adam operation for operation:
bit-identical to torch whenever the clip coefficient clamps to 1, fp32-rounding-close otherwise (the total norm is
summed in a different order).

The optimizer's own state is used and kept in torch's *capturable* layout (``state['step']`` a float32 device
scalar; ``param_group['capturable'] = True``), so ``optimizer.state_dict()`` / ``load_state_dict`` / a plain
``optimizer.step()`` keep working, and the ``GradScaler`` keeps its own ``_scale`` / ``_growth_tracker`` tensors,
updated in place.  CUDA fp32 parameters only; no CPU fallback.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional

import torch

from . import _lib


class FusedAdamStep:
    """``FusedAdamStep(optimizer, scaler, max_grad_norm).step()`` == ``AMPContext.step_optimizer`` (amp_utils.py:29-41)."""

    def __init__(self, optimizer: torch.optim.Optimizer, scaler: Optional["torch.amp.GradScaler"] = None,
                 max_grad_norm: Optional[float] = None):
        if not isinstance(optimizer, torch.optim.Adam) or isinstance(optimizer, torch.optim.AdamW):
            raise TypeError("FusedAdamStep drives a torch.optim.Adam (what sched_optim.get_optimizer builds)")
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("one parameter group (the reference builds Adam(model.parameters(), ...))")
        g = optimizer.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("differentiable") or g.get("decoupled_weight_decay"):
            raise NotImplementedError("amsgrad / maximize / differentiable / decoupled weight decay are not supported")
        if isinstance(g["lr"], torch.Tensor) or any(isinstance(b, torch.Tensor) for b in g["betas"]):
            raise NotImplementedError("tensor lr / betas are not supported")
        self.optimizer, self.scaler, self.max_grad_norm = optimizer, scaler, max_grad_norm
        self._key = None
        self._dev = None
        self._scratch = None
        g["capturable"] = True            # state['step'] lives on the device; torch's own step() stays usable

    # -- tables ------------------------------------------------------------------------------------
    def _prepare(self, params):
        opt = self.optimizer
        dev = params[0].device
        for p in params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.device == dev):
                raise RuntimeError("FusedAdamStep: parameters must be contiguous fp32 CUDA tensors on one device "
                                   "(the B200 path has no CPU fallback)")
            gr = p.grad
            if gr.is_sparse or gr.dtype != torch.float32 or gr.device != dev:
                raise RuntimeError("FusedAdamStep: gradients must be dense fp32 tensors on the parameters' device")
            if not gr.is_contiguous():
                p.grad = gr.contiguous()
            st = opt.state[p]
            if len(st) == 0:                                          # Adam._init_group, capturable layout
                st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif not st["step"].is_cuda:                              # state made by torch's default path / a checkpoint
                st["step"] = st["step"].to(device=dev, dtype=torch.float32)
            if not (st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
                raise RuntimeError("FusedAdamStep: optimizer state must be contiguous")
        key = tuple((p.data_ptr(), p.grad.data_ptr(), opt.state[p]["exp_avg"].data_ptr(),
                     opt.state[p]["exp_avg_sq"].data_ptr(), opt.state[p]["step"].data_ptr(), p.numel()) for p in params)
        if key == self._key:
            return
        lib = _lib.load()
        n = len(params)
        if self._key is None or tuple(k[5] for k in key) != tuple(k[5] for k in self._key) or dev != self._dev:
            chunk = lib.pnce_multi_chunk_elems()
            ct, cs = [], []
            for t, k in enumerate(key):
                for e in range(0, k[5], chunk):
                    ct.append(t)
                    cs.append(e)
            self._n_chunks = len(ct)
            self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
            self._chunk_start = torch.tensor(cs, dtype=torch.int64, device=dev)
            self._numel = torch.tensor([k[5] for k in key], dtype=torch.int64, device=dev)
            self._scratch = torch.zeros(lib.pnce_amp_adam_scratch_floats(self._n_chunks), dtype=torch.float32, device=dev)
        # one (5, n) pointer table: param, grad, exp_avg, exp_avg_sq, step
        self._ptrs = torch.tensor([[k[c] for k in key] for c in range(5)], dtype=torch.int64).to(dev)
        self._key, self._dev, self._n = key, dev, n

    # -- the call -----------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self):
        opt, scaler = self.optimizer, self.scaler
        g = opt.param_groups[0]
        params = [p for p in g["params"] if p.grad is not None]         # amp_utils.py:35 / Adam._init_group
        use_scaler = scaler is not None and scaler.is_enabled()
        if not params:
            if use_scaler:
                raise AssertionError("No inf checks were recorded for this optimizer.")   # what GradScaler.step raises
            return
        self._prepare(params)
        scale_ptr = tracker_ptr = None
        gf = bf = 1.0
        gi = 1
        if use_scaler:
            if scaler._scale is None:
                raise AssertionError("Attempted step but _scale is None.  This may indicate your script did not use "
                                     "scaler.scale(loss or outputs) earlier in the iteration.")
            if scaler._scale.device != self._dev:
                raise RuntimeError("GradScaler and parameters live on different devices")
            scale_ptr, tracker_ptr = scaler._scale.data_ptr(), scaler._growth_tracker.data_ptr()
            gf, bf, gi = scaler.get_growth_factor(), scaler.get_backoff_factor(), scaler.get_growth_interval()
        lib = _lib.load()
        row = self._ptrs.data_ptr()
        stride = self._n * 8
        beta1, beta2 = g["betas"]
        mx = -1.0 if self.max_grad_norm is None else float(self.max_grad_norm)
        _lib.check(lib.pnce_amp_adam_step(
            row, row + stride, row + 2 * stride, row + 3 * stride, row + 4 * stride, self._numel.data_ptr(),
            self._chunk_tensor.data_ptr(), self._chunk_start.data_ptr(), self._n_chunks, self._n,
            scale_ptr, tracker_ptr, ctypes.c_float(gf), ctypes.c_float(bf), int(gi), ctypes.c_float(mx),
            float(g["lr"]), float(beta1), float(beta2), float(g["eps"]), float(g["weight_decay"]),
            self._scratch.data_ptr(), torch.cuda.current_stream(self._dev).cuda_stream), "pnce_amp_adam_step")
        # parameters (and gradients: unscaled + clipped in place) were written through raw pointers -- bump the
        # version counters like an in-place torch op would (feature_reuse's staleness guard reads them)
        torch._C._increment_version(params)
        torch._C._increment_version([p.grad for p in params])

    def last_total_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step (after unscaling, before clipping) -- a device scalar, no sync."""
        return self._scratch[1]


_STEPPERS = weakref.WeakKeyDictionary()


def amp_step_optimizer(amp_ctx, optimizer, max_grad_norm=None):
    """``AMPContext.step_optimizer`` (amp_utils.py:29-41) as a free function / a method replacement:
    ``AMPContext.step_optimizer = amp_step_optimizer``.  ``amp_ctx`` only needs ``.scaler``."""
    st = _STEPPERS.get(optimizer)
    if st is None or st.scaler is not amp_ctx.scaler:
        st = _STEPPERS[optimizer] = FusedAdamStep(optimizer, amp_ctx.scaler, max_grad_norm)
    st.max_grad_norm = max_grad_norm
    st.step()
