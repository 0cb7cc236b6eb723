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
        # The kernel updates the WHOLE flat arena.  Parameters that were not passed, or have requires_grad=False
        # (transfer learning, utils.load_model_components), are kept fixed by zeroing their gradient ranges before every
        # step: with zero gradient and zero moments the Adam update is exactly 0 -- provided there is no weight decay.
        mine = {id(p) for p in self.param_groups[0]["params"] if p.requires_grad}
        ranges = []
        for n, p in owner.named_parameters():
            if n in owner.arena.offsets and id(p) not in mine:
                lo = owner.arena.offsets[n]
                hi = lo + p.numel()
                if ranges and ranges[-1][1] == lo:
                    ranges[-1][1] = hi
                else:
                    ranges.append([lo, hi])
        if ranges and weight_decay != 0:
            raise ValueError("FusedAdam with frozen / omitted parameters needs weight_decay=0 (coupled L2 would still move "
                             "them); use torch.optim.Adam for that combination")
        self._inactive = [tuple(r) for r in ranges]
        self._owner = owner
        self._m = None
        self._v = None
        self.grad_scale = 1.0

    def _moments(self, like):
        if self._m is None or self._m.device != like.device:
            self._m = torch.zeros_like(like)
            self._v = torch.zeros_like(like)
        return self._m, self._v

    def fused_step(self, engine, advance_rng: bool = False):
        g = self.param_groups[0]
        m, v = self._moments(engine.params)
        for lo, hi in self._inactive:
            engine.ops.fill_f32(engine.grads[lo:hi], 0.0)
        engine.adam(m, v, g["lr"], g["betas"], g["eps"], g["weight_decay"], self.grad_scale, advance_rng=advance_rng)
        self._opt_called = True        # what torch's step() wrapper records; lr schedulers check it before their own step

    # moments and the step counter are flat tensors outside torch's per-parameter ``state``: carry them explicitly
    def state_dict(self):
        sd = super().state_dict()
        model = self._owner
        sd["csi_flat"] = {"exp_avg": None if self._m is None else self._m.detach().cpu().clone(),
                          "exp_avg_sq": None if self._v is None else self._v.detach().cpu().clone(),
                          "step": None if model._opt_step is None else int(model._opt_step.item())}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("csi_flat", None)
        super().load_state_dict(state_dict)
        if flat is not None:
            model = self._owner
            dev = model.flat_params.device
            if flat["exp_avg"] is not None:
                self._m, self._v = flat["exp_avg"].to(dev).clone(), flat["exp_avg_sq"].to(dev).clone()
            if flat["step"] is not None:
                model._counters(dev)[1].fill_(int(flat["step"]))

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
