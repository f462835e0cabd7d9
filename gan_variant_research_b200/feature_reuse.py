"""Encoder-feature reuse.

The reference's G step runs the generator three times on the way to the PatchNCE loss:
``fake = generator(photos)`` (training/train_cutpp.py:270), then inside ``compute_patchnce_loss``
``generator.get_feature_layers(photos, nce_layers)`` under ``no_grad`` (losses/patchnce_cut.py:138-139)
and ``generator.get_feature_layers(fake, nce_layers)`` with grad (:145).  The second of these recomputes,
value for value, activations the first one has just produced: ``get_feature_layers``
(models/generator_resnet_attn.py:190-235) walks the same ``initial`` / ``downsample`` / ``res_blocks`` /
``upsample`` stacks as ``forward`` (:165-188) and only differs in returning the intermediate maps.

``EncoderFeatureCache`` taps those maps with forward hooks while ``generator(x)`` runs and hands them to
``compute_patchnce_loss`` when it is asked for the features of *the same tensor*: one generator pass per
G step disappears and the gather reads maps that were written a moment ago.  The training loop is not
edited: ``enable_encoder_feature_reuse(generator, nce_layers)`` once, before the loop.

Logical layer numbering (generator_resnet_attn.py:205-233): 0 = output of ``initial``; then one index
per ``nn.ReLU`` of ``downsample``; one per residual block; one per ``nn.ReLU`` of ``upsample``.
Ids past the last one are silently absent, as in the reference.

A hit requires: the same tensor object, unchanged since the capture (``_version``), parameters unchanged
since the capture, the same train/eval mode and the same autocast state -- otherwise the lookup misses
and the caller recomputes, exactly as the reference does.  Captured maps are handed out detached (the
reference detaches the source features, patchnce_cut.py:142) and released on the first lookup.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

_ATTR = "_pnce_encoder_feature_cache"


def logical_layer_modules(generator: nn.Module) -> List[nn.Module]:
    """Module whose *output* is logical layer i of ``get_feature_layers``
    (generator_resnet_attn.py:205-233), for i = 0, 1, ..."""
    for name in ("initial", "downsample", "res_blocks", "upsample"):
        if not hasattr(generator, name):
            raise TypeError(f"encoder-feature reuse needs the reference generator's layout "
                            f"(initial / downsample / res_blocks / upsample); '{name}' is missing")
    mods: List[nn.Module] = [generator.initial]                                        # :208-211
    mods += [m for m in generator.downsample if isinstance(m, nn.ReLU)]                # :214-219
    mods += list(generator.res_blocks)                                                 # :222-226
    mods += [m for m in generator.upsample if isinstance(m, nn.ReLU)]                  # :229-234
    return mods


def _autocast_state(device_type: str):
    if not torch.is_autocast_enabled(device_type):
        return (False, None)
    return (True, torch.get_autocast_dtype(device_type))


class EncoderFeatureCache:
    """Forward-hook tap of the maps ``get_feature_layers`` would return for the last ``generator(x)``."""

    def __init__(self, generator: nn.Module, nce_layers: Sequence[int]):
        self.generator = generator
        self.nce_layers = sorted(set(int(i) for i in nce_layers))
        mods = logical_layer_modules(generator)
        self._n_logical = len(mods)
        self._handles = []
        self._inside = False
        self._key = None
        self._maps: Dict[int, torch.Tensor] = {}
        self._input: Optional[torch.Tensor] = None
        self.hits = 0
        self.misses = 0
        self._handles.append(generator.register_forward_pre_hook(self._enter))
        self._handles.append(generator.register_forward_hook(self._leave))
        for idx in self.nce_layers:
            if idx < self._n_logical:
                self._handles.append(mods[idx].register_forward_hook(self._make_tap(idx)))

    # -- capture -------------------------------------------------------------------------------
    def _params_version(self) -> int:
        return sum(p._version for p in self.generator.parameters())

    def _state_key(self, x: torch.Tensor):
        return (x._version, self._params_version(), self.generator.training, _autocast_state(x.device.type))

    def _enter(self, module, args):
        # only generator.__call__ opens a capture: get_feature_layers calls the stacks directly and must
        # not overwrite the maps of the forward pass it is about to be served from
        self._maps = {}
        self._input = args[0] if args and isinstance(args[0], torch.Tensor) else None
        self._key = self._state_key(self._input) if self._input is not None else None
        self._inside = self._input is not None

    def _make_tap(self, idx: int):
        def tap(module, args, output):
            if self._inside:
                self._maps[idx] = output
        return tap

    def _leave(self, module, args, output):
        self._inside = False

    # -- lookup --------------------------------------------------------------------------------
    def lookup(self, x: torch.Tensor, layer_ids: Optional[Sequence[int]]) -> Optional[List[torch.Tensor]]:
        """The list ``generator.get_feature_layers(x, layer_ids)`` would return, detached, if ``x`` is the
        tensor the generator was last called on and nothing relevant changed since; else ``None``."""
        if layer_ids is None:
            layer_ids = [0, 4, 8, 12, 16]                                              # :201-202
        want = sorted(set(int(i) for i in layer_ids if int(i) < self._n_logical))
        ok = (self._input is not None and x is self._input and not self._inside
              and self._key == self._state_key(x) and all(i in self._maps for i in want))
        if not ok:
            self.misses += 1
            return None
        feats = [self._maps[i].detach() for i in want]     # ascending logical order, as the reference appends
        self._maps, self._input, self._key = {}, None, None
        self.hits += 1
        return feats

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        self._maps, self._input, self._key = {}, None, None
        if getattr(self.generator, _ATTR, None) is self:
            delattr(self.generator, _ATTR)


def enable_encoder_feature_reuse(generator: nn.Module, nce_layers: Sequence[int]) -> EncoderFeatureCache:
    """Attach an ``EncoderFeatureCache`` to ``generator``; ``compute_patchnce_loss`` finds it there.
    Idempotent per generator (a second call replaces the first cache)."""
    old = getattr(generator, _ATTR, None)
    if old is not None:
        old.remove()
    cache = EncoderFeatureCache(generator, nce_layers)
    object.__setattr__(generator, _ATTR, cache)
    return cache


def cached_source_features(generator, src_images, nce_layers) -> Optional[List[torch.Tensor]]:
    cache = getattr(generator, _ATTR, None)
    if cache is None:
        return None
    return cache.lookup(src_images, nce_layers)
