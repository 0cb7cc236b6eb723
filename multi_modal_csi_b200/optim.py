"""Fused Adam over the flat parameter arena of a THAT model.

Replaces ``torch.optim.Adam(model.parameters(), lr=..., weight_decay=...)`` at
benchmark/wifi_csi/model/that.py:395-397: same update rule (coupled L2: grad += weight_decay * param; bias
correction; eps added after the square root), same constructor arguments, but one kernel launch over the
118 tensors instead of a per-tensor loop, with its step counter on the device so the whole train step can be
replayed as a CUDA graph.
"""
from __future__ import annotations

import torch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam takes the parameters of one THAT model as a single group")
        owners = {id(getattr(p, "_csi_owner", lambda: None)()) for p in self.param_groups[0]["params"]}
        owner = getattr(self.param_groups[0]["params"][0], "_csi_owner", lambda: None)()
        if owner is None or len(owners) != 1:
            raise ValueError("FusedAdam needs the parameters of a multi_modal_csi_b200.THAT model")
        self._owner = owner
        self._m = None
        self._v = None
        self.grad_scale = 1.0

    def _moments(self, like):
        if self._m is None or self._m.device != like.device:
            self._m = torch.zeros_like(like)
            self._v = torch.zeros_like(like)
        return self._m, self._v

    def fused_step(self, engine):
        g = self.param_groups[0]
        m, v = self._moments(engine.params)
        engine.adam(m, v, g["lr"], g["betas"], g["eps"], g["weight_decay"], self.grad_scale)
        self._opt_called = True        # what torch's step() wrapper records; lr schedulers check it before their own step

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self._owner
        eng = model._engine
        if eng is None:
            raise RuntimeError("FusedAdam.step() before any forward/backward of the model")
        self.fused_step(eng)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # gradients live in the flat arena; the next backward overwrites it
        for p in self.param_groups[0]["params"]:
            p.grad = None
