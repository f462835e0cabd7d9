"""Parameter EMA in one launch: drop-in for ``EMA`` of the reference's
``GAN_Variant1/utils/io_ckpt.py:9-53`` -- same constructor, ``update`` / ``apply_shadow`` / ``restore`` /
``state_dict`` / ``load_state_dict``, same ``{'decay', 'shadow': {name: tensor}}`` checkpoint layout
(``generate_folder.py:130-135`` reads ``ckpt['ema_G']['shadow']``), bit-identical values.

The reference walks ``named_parameters()`` in Python and per tensor launches two multiplies, an add and a
clone (plus two allocations); here the shadow values live in ONE flat fp32 buffer (``shadow[name]`` are views of
it) and ``update()`` is a single multi-tensor launch of libpnce (``pnce_multi_axpby``) over a device table built
once -- parameter storage does not move during training.  HBM-bound: 12 bytes per parameter.
CUDA parameters only (no CPU fallback)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class EMA:
    def __init__(self, model: torch.nn.Module, decay: float = 0.999):
        self.model = model
        self.decay = decay
        self.backup = {}
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]      # :19-21
        if not named:
            raise RuntimeError("EMA: the model has no trainable parameters")
        for n, p in named:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"EMA: parameter {n} must be a contiguous fp32 CUDA tensor "
                                   "(the B200 path has no CPU fallback; use the reference's EMA on CPU)")
        self._names = [n for n, _ in named]
        self._params = [p for _, p in named]
        dev = self._params[0].device
        self._dev = dev
        sizes = [p.numel() for p in self._params]
        # every view starts on a 16-byte boundary so the kernel can use 128-bit accesses
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (s + 3) // 4 * 4
        self._flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._backup_flat = None
        self.shadow = {n: self._flat[o:o + s].view_as(p) for n, o, s, p in zip(self._names, offs, sizes, self._params)}
        self._offs, self._sizes = offs, sizes
        lib = _lib.load()
        chunk = lib.pnce_multi_chunk_elems()
        ct, cs = [], []
        for t, s in enumerate(sizes):
            for e in range(0, s, chunk):
                ct.append(t)
                cs.append(e)
        self._n_chunks = len(ct)
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self._chunk_start = torch.tensor(cs, dtype=torch.int64, device=dev)
        self._numel = torch.tensor(sizes, dtype=torch.int64, device=dev)
        self._shadow_ptrs = torch.tensor([v.data_ptr() for v in self.shadow.values()], dtype=torch.int64, device=dev)
        self._param_ptrs = None
        self._refresh_param_table()
        self._launch(self._shadow_ptrs, self._param_ptrs, 0.0, 0.0, 1)                   # shadow = param.clone()

    def _refresh_param_table(self):
        ptrs = [p.data_ptr() for p in self._params]
        if self._param_ptrs is None or ptrs != self._param_ptr_list:
            # storage moved (.to(memory_format=...), .to(device), re-created parameters): the kernel walks raw storage in the
            # shadow's contiguous order, so a channels-last or otherwise strided parameter must fail loudly, not average
            # permuted values
            for n, p in zip(self._names, self._params):
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != self._dev:
                    raise RuntimeError(f"EMA: parameter {n} is no longer a contiguous fp32 CUDA tensor on {self._dev}")
            self._param_ptr_list = ptrs
            self._param_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self._dev)

    def _launch(self, dst_table, src_table, a, b, mode):
        lib = _lib.load()
        st = torch.cuda.current_stream(self._dev).cuda_stream
        _lib.check(lib.pnce_multi_axpby(dst_table.data_ptr(), src_table.data_ptr(), self._numel.data_ptr(),
                                        self._chunk_tensor.data_ptr(), self._chunk_start.data_ptr(), self._n_chunks,
                                        ctypes.c_float(a), ctypes.c_float(b), mode, st), "pnce_multi_axpby")

    def update(self):
        """shadow = (1 - decay) * param + decay * shadow for every trainable parameter -- :23-29."""
        self._refresh_param_table()          # a cheap list compare; storage only moves if the user re-creates parameters
        self._launch(self._shadow_ptrs, self._param_ptrs, 1.0 - self.decay, self.decay, 0)

    def apply_shadow(self):
        """Put the EMA values into the model (for evaluation), keeping a backup -- :31-36."""
        self._refresh_param_table()
        if self._backup_flat is None:
            self._backup_flat = torch.empty_like(self._flat)
            self._backup_ptrs = torch.tensor([self._backup_flat[o:o + s].data_ptr() for o, s in zip(self._offs, self._sizes)],
                                             dtype=torch.int64, device=self._dev)
        self._launch(self._backup_ptrs, self._param_ptrs, 0.0, 0.0, 1)
        self.backup = {n: self._backup_flat[o:o + s].view_as(p)
                       for n, o, s, p in zip(self._names, self._offs, self._sizes, self._params)}
        self._launch(self._param_ptrs, self._shadow_ptrs, 0.0, 0.0, 1)
        # the kernel wrote the parameters through raw pointers: tell autograd's version counters, which is what
        # anything that caches values derived from the parameters checks (feature_reuse._params_version)
        torch._C._increment_version(self._params)

    def restore(self):
        """Put the original parameters back -- :38-43."""
        if not self.backup:
            raise KeyError("restore() without apply_shadow()")      # the reference raises KeyError on backup[name]
        self._refresh_param_table()
        self._launch(self._param_ptrs, self._backup_ptrs, 0.0, 0.0, 1)
        torch._C._increment_version(self._params)
        self.backup = {}

    def state_dict(self):
        return {"decay": self.decay, "shadow": self.shadow}          # :45-49

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]                              # :51-53 (values are copied into the flat buffer)
        for n, v in state_dict["shadow"].items():
            self.shadow[n].copy_(v.to(self._dev, torch.float32))
