"""Base class of the models of this package: an ``nn.Module`` whose trainable parameters are views of ONE flat fp32
arena (and their ``.grad`` views of a second one), so that the optimizer is one fused kernel (``optim.FusedAdam``) and
the data-parallel exchange a contiguous all-reduce (``parallel.GradSync``).  The module tree only exists to reproduce the
reference's ``state_dict`` keys.  Subclasses set ``self.arena`` (layout.Arena), ``self._flat`` / ``self._gflat`` and
register their tensors with ``_register``; ``FROZEN_PARAMS`` names non-trainable parameters kept outside the arena."""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import layout as LY


class _Node(torch.nn.Module):
    """Parameter container; the module tree only exists to reproduce the reference's state_dict keys."""


class ArenaModule(torch.nn.Module):
    FROZEN_PARAMS = ()

    @property
    def rng_seed(self) -> int:
        return self._rng_seed

    @rng_seed.setter
    def rng_seed(self, seed: int):
        self._rng_seed = int(seed)
        if self._rng is not None:
            self._rng[0] = self._rng_seed

    def _counters(self, device):
        if self._rng is None:
            self._rng = torch.tensor([self._rng_seed, 0], dtype=torch.int64, device=device)
            self._opt_step = torch.ones(1, dtype=torch.int64, device=device)
        elif self._rng.device != device:
            self._rng, self._opt_step = self._rng.to(device), self._opt_step.to(device)
        return self._rng, self._opt_step

    # ------------------------------------------------------------------ module tree
    def _register(self, name, tensor, is_buffer):
        parts = name.split(".")
        node = self
        for i, part in enumerate(parts[:-1]):
            nxt = parts[i + 1]
            if part.isdigit():
                idx = int(part)
                assert isinstance(node, torch.nn.ModuleList)
                while len(node) <= idx:
                    node.append(torch.nn.ModuleList() if nxt.isdigit() else _Node())
                node = node[idx]
            else:
                if part not in node._modules:
                    node.add_module(part, torch.nn.ModuleList() if nxt.isdigit() else _Node())
                node = node._modules[part]
        if is_buffer:
            node.register_buffer(parts[-1], tensor)
        else:
            node.register_parameter(parts[-1], tensor)

    def _named(self):
        return OrderedDict(self.named_parameters())

    def _apply(self, fn, recurse=True):
        """``.to(device)`` / ``.cuda()``: move the two flat arenas and re-create the parameter views."""
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise TypeError("THAT master weights are fp32; choose the compute type with act_dtype")
        new_g = fn(self._gflat)
        object.__setattr__(self, "_flat", new_flat)
        object.__setattr__(self, "_gflat", new_g)
        for name, p in self._named().items():
            had_grad = p.grad is not None
            if name in self.FROZEN_PARAMS:
                p.data = fn(p.data)
            else:
                off, shape = self.arena.offsets[name], self.arena.shapes[name]
                p.data = new_flat[off:off + LY.numel(shape)].view(shape)
                p.grad = new_g[off:off + LY.numel(shape)].view(shape) if had_grad else None
        for mod in self.modules():
            for k, b in mod._buffers.items():
                if b is not None:
                    mod._buffers[k] = fn(b)
        self._engine = None
        return self

    def _attach_grads(self):
        for name, p in self.named_parameters():
            if p.requires_grad and p.grad is None:
                off, shape = self.arena.offsets[name], self.arena.shapes[name]
                p.grad = self._gflat[off:off + LY.numel(shape)].view(shape)

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_grads(self) -> torch.Tensor:
        return self._gflat
