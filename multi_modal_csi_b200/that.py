"""Drop-in for benchmark/wifi_csi/model/that.py: ``THAT(var_x_shape, var_y_shape)`` and ``run_that(...)``.

The module keeps the reference's parameter names, shapes, registration order and initialisation calls
(that.py:31-56,100-139,180-245), so ``torch.random.manual_seed(r + 39)`` produces the same initial weights and
``state_dict()`` / ``.pth`` files interchange with the reference in both directions.  All parameters are views of
one flat fp32 arena (and their ``.grad`` of a second one): that is what lets the optimizer be one fused kernel
and the data-parallel all-reduce run over a few contiguous buckets.

``forward`` runs the hand-written sm_100a kernels of ``libcsi_that.so`` through ``THATEngine``; there is no
PyTorch-operator or CPU fallback: on a non-CUDA device, or without the library, it raises.
"""
from __future__ import annotations

import math
import os
import weakref
from collections import OrderedDict
from typing import Optional

import torch

from . import layout as LY
from .arena_module import ArenaModule, _Node
from .engine import THATEngine


def _kaiming_conv(shape):
    """torch.nn.Conv1d / Linear.reset_parameters(): kaiming_uniform_(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias."""
    w = torch.empty(*shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = LY.numel(shape[1:])
    bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
    b = torch.empty(shape[0]).uniform_(-bound, bound)
    return w, b


def _initial_values(g: LY.ModelGeom):
    """Initial tensors, drawn in the order the reference's constructors consume the CPU RNG."""
    vals = OrderedDict()
    bufs = OrderedDict()
    L, F = g.left.L, g.F
    K = LY.NUM_GAUSS
    if g.heads > 1:                                    # that_multi_head.py:194-196: the heads are constructed first
        for wn, bn in g.output_names():
            vals[wn], vals[bn] = _kaiming_conv((g.out, LY.FEAT))
    emb = torch.zeros(K, F)
    torch.nn.init.xavier_uniform_(emb)                                               # that.py:43-45
    vals["layer_left_gaussian.var_embedding"] = emb
    vals["layer_left_gaussian.var_position"] = torch.arange(0.0, L).unsqueeze(1).repeat(1, K)   # that.py:48
    vals["layer_left_gaussian.var_mu"] = torch.arange(0.0, L, L / K).unsqueeze(0)    # that.py:52
    vals["layer_left_gaussian.var_sigma"] = torch.tensor([50.0] * K).unsqueeze(0)    # that.py:56

    def enc(p, d, kernels):
        vals[p + "layer_norm_0.weight"] = torch.ones(d)
        vals[p + "layer_norm_0.bias"] = torch.zeros(d)
        # nn.MultiheadAttention.__init__: out_proj Linear is reset first, then _reset_parameters()
        ow, _ob = _kaiming_conv((d, d))
        iw = torch.empty(3 * d, d)
        torch.nn.init.xavier_uniform_(iw)
        vals[p + "layer_attention.in_proj_weight"] = iw
        vals[p + "layer_attention.in_proj_bias"] = torch.zeros(3 * d)
        vals[p + "layer_attention.out_proj.weight"] = ow
        vals[p + "layer_attention.out_proj.bias"] = torch.zeros(d)
        vals[p + "layer_norm_1.weight"] = torch.ones(d)
        vals[p + "layer_norm_1.bias"] = torch.zeros(d)
        for j, k in enumerate(kernels):
            w, b = _kaiming_conv((d, d, k))
            vals[f"{p}layer_cnn.{j}.0.weight"] = w
            vals[f"{p}layer_cnn.{j}.0.bias"] = b
            vals[f"{p}layer_cnn.{j}.1.weight"] = torch.ones(d)
            vals[f"{p}layer_cnn.{j}.1.bias"] = torch.zeros(d)
            bufs[f"{p}layer_cnn.{j}.1.running_mean"] = torch.zeros(d)
            bufs[f"{p}layer_cnn.{j}.1.running_var"] = torch.ones(d)
            bufs[f"{p}layer_cnn.{j}.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for s in g.streams:
        for e in range(s.n_enc):
            enc(s.prefix(e), s.d, s.kernels)
        vals[f"layer_{s.name}_norm.weight"] = torch.ones(s.d)
        vals[f"layer_{s.name}_norm.bias"] = torch.zeros(s.d)
        for j, k in enumerate(s.head_k):
            w, b = _kaiming_conv((s.head_n, s.d, k))
            vals[f"layer_{s.name}_cnn_{j}.weight"] = w
            vals[f"layer_{s.name}_cnn_{j}.bias"] = b
    if g.heads == 1:
        w, b = _kaiming_conv((g.out, LY.FEAT))
        vals["layer_output.weight"] = w
        vals["layer_output.bias"] = b
    return vals, bufs


class _THATFunction(torch.autograd.Function):
    """Autograd bridge for the reference-style loop ``loss(model(x), y).backward()`` (train.py:96-100)."""

    @staticmethod
    def forward(ctx, x, anchor, model):
        B = x.shape[0]
        eng = model._engine_for(B)
        eng.repack()
        eng.begin_train_forward()
        logits = eng.forward(x, B, training=True, dropout=model.dropout_enabled, augment=False)
        ctx.model, ctx.B = model, B
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        model, B = ctx.model, ctx.B
        eng = model._engine
        params = [p for p in model.parameters() if p.requires_grad]
        fresh = all(p.grad is None for p in params)
        eng.backward(dlogits.contiguous().float(), B, dropout=model.dropout_enabled, zero_grads=fresh)
        eng.end_train_step()             # the dropout masks of this forward/backward pair are spent: next Philox step
        model._attach_grads()
        return None, None, None


class THAT(ArenaModule):
    """``THAT(var_x_shape, var_y_shape)``: var_x_shape[-2:] = (T, F), var_y_shape[-1] = out (that.py:183-192)."""

    NUM_OUTPUT_HEADS = 1
    FROZEN_PARAMS = LY.FROZEN

    def __init__(self, var_x_shape, var_y_shape, act_dtype: Optional[str] = None, max_batch: Optional[int] = None):
        super().__init__()
        F, T, out = int(var_x_shape[-1]), int(var_x_shape[-2]), int(var_y_shape[-1])
        self.geom = LY.ModelGeom(T, F, out, self.NUM_OUTPUT_HEADS)
        self.specs = LY.parameter_specs(self.geom)
        self.arena = LY.build_arena(self.specs)
        self.act_dtype = {None: torch.bfloat16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
                          "fp32": torch.float32, "float32": torch.float32}[act_dtype]
        self.max_batch = max_batch
        self.dropout_enabled = True          # parity tests switch the Dropout layers off (p = 0)
        # Philox {seed, step} and the 1-based Adam step live on the MODEL (device tensors, created with the first
        # engine): an engine rebuilt for a larger batch, or after .to() / configure(), keeps counting where the old one was
        self._rng_seed = int(torch.initial_seed() & 0x7FFFFFFF)
        self._rng = None
        self._opt_step = None
        self._engine: Optional[THATEngine] = None
        self._ops_override = None            # tests only: inject the torch mirror of the kernels
        # fused_train_step replays forward+loss+backward as one CUDA graph (CSI_NO_GRAPH=1: eager launches, for profilers)
        self.use_cuda_graph = os.environ.get("CSI_NO_GRAPH", "0") != "1"
        self._eager_steps = 0
        vals, bufs = _initial_values(self.geom)
        flat = torch.zeros(self.arena.size)
        object.__setattr__(self, "_flat", flat)
        object.__setattr__(self, "_gflat", torch.zeros(self.arena.size))
        for name, shape in self.specs.items():
            frozen = name in LY.FROZEN
            if frozen:
                data = vals[name].clone()
            else:
                off = self.arena.offsets[name]
                data = flat[off:off + LY.numel(shape)].view(shape)
                data.copy_(vals[name])
            p = torch.nn.Parameter(data, requires_grad=not frozen)
            p._csi_owner = weakref.ref(self)
            p._csi_name = name
            self._register(name, p, is_buffer=False)
        for name, val in bufs.items():
            self._register(name, val, is_buffer=True)

    # ------------------------------------------------------------------ engine
    def configure(self, act_dtype: Optional[str] = None, max_batch: Optional[int] = None):
        """Optional preset keys ``nn.dtype`` / batch size feed this; defaults reproduce the reference call."""
        if act_dtype is not None:
            self.act_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[act_dtype]
        if max_batch is not None:
            self.max_batch = max_batch
        self._engine = None
        return self

    def _engine_for(self, B: int) -> THATEngine:
        eng = self._engine
        if eng is None or eng.B < B:
            if self._flat.device.type != "cuda" and self._ops_override is None:
                raise RuntimeError("multi_modal_csi_b200.THAT runs only on a CUDA (sm_100a) device: "
                                   "call .to('cuda') first; there is no CPU path")
            mb = max(B, self.max_batch or 0)
            bn = OrderedDict((k, v) for k, v in self.named_buffers())
            frozen = {k: p.data.reshape(-1) for k, p in self.named_parameters() if k in LY.FROZEN}
            rng, opt_step = self._counters(self._flat.device)
            eng = THATEngine(self.geom, mb, self._flat, self._gflat, self.arena, bn, frozen,
                             act_dtype=self.act_dtype, ops=self._ops_override, rng=rng, opt_step=opt_step)
            if self._engine is not None:
                eng.rng_used = self._engine.rng_used
            self._engine = eng
            self._eager_steps = 0
        return eng

    # ------------------------------------------------------------------ forward
    def forward(self, var_input: torch.Tensor) -> torch.Tensor:
        """float32 [B, T, F] -> float32 logits [B, out]  (that.py:249-302)."""
        x = var_input
        if x.dim() != 3 or x.shape[1] != self.geom.T or x.shape[2] != self.geom.F:
            raise ValueError(f"expected input [B,{self.geom.T},{self.geom.F}], got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        if x.device != self._flat.device:
            raise RuntimeError(f"input on {x.device} but model on {self._flat.device}")
        if self.training and torch.is_grad_enabled():
            anchor = torch.zeros((), device=x.device, requires_grad=True)   # ties the graph to backward()
            return _THATFunction.apply(x, anchor, self)
        return self._forward_nograd(x, self.training)

    def _forward_nograd(self, x, training):
        N = x.shape[0]
        eng = self._engine_for(min(N, self.max_batch or 256))
        eng.repack()
        outs = []
        for i in range(0, N, eng.B):
            xb = x[i:i + eng.B]
            if training:
                eng.begin_train_forward()
            outs.append(eng.forward(xb, xb.shape[0], training=training, dropout=self.dropout_enabled).clone())
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    # ------------------------------------------------------------------ prediction rule on the device
    def predict_counts(self, logits: torch.Tensor, users: int = 6, threshold: float = 0.5) -> torch.Tensor:
        """Per-class people counts [N, classes] (int32) from logits [N, users*classes] with the reference rule
        (utils.py:147-183: per-user arg-max kept iff its sigmoid exceeds the threshold), computed on the GPU."""
        z = logits.float().contiguous()
        n, classes = z.shape[0], z.shape[1] // users
        counts = torch.zeros(n, classes, dtype=torch.int32, device=z.device)
        self._engine_for(1).ops.predict_counts(z, n, users, classes, threshold, counts)
        return counts

    # ------------------------------------------------------------------ fused train step
    def fused_train_step(self, x, y, optimizer, pos_weight: float = 4.0, augment: bool = True,
                         grad_hook=None, offs=None, lens=None, use_graph=None, loss_kind: str = "bce"):
        """augmentation + forward + BCE + backward + Adam as one launch sequence (train.py:84-101).

        x: fp32 [B,T,F] on the device (or a packed ragged arena with offs/lens); y: [B, ...] labels.
        ``grad_hook(engine)`` runs between backward and the optimizer (data-parallel all-reduce).
        Returns (loss 1-element tensor, logits [B,out]); both are views of static buffers."""
        if loss_kind not in ("bce", "smooth_l1", "perm_ce"):
            raise ValueError(f"unsupported fused loss {loss_kind!r}")
        B = y.shape[0]
        eng = self._engine_for(B)
        eng.loss_kind = loss_kind
        yf = y.reshape(B, -1)
        if yf.dtype != torch.float32:
            yf = yf.float()
        eng.begin_train_forward()
        eng.forward_input(x, B, True, augment, offs, lens)
        eng.ops.copy_f32(eng.y_static, yf.contiguous(), yf.numel())     # rows of y_static and yf have the same pitch
        eng.ensure_packed()                             # operand copies: packed after the previous optimizer step (side stream)
        graph = self.use_cuda_graph if use_graph is None else use_graph
        run = eng.train_body_graph if (graph and x.is_cuda and self._eager_steps >= 1) else eng.train_body
        overlap = grad_hook is not None and hasattr(grad_hook, "start_bucket")
        if overlap and getattr(grad_hook, "one_graph", False) and x.is_cuda and eng.g.left.n_enc > 1:
            # data parallel, one launch sequence: both all-reduces are issued from inside backward (captured into the step
            # graph), the first one waits only for the streams that feed its bucket
            run(B, pos_weight, self.dropout_enabled, 0, grad_hook)
        elif overlap:
            # data parallel: all-reduce the first gradient bucket (everything but the Gaussian encoding and left
            # encoder 0: ~3/4 of the bytes) on a side stream while the rest of the left stream's backward runs
            (lo1, hi1), (lo2, hi2) = eng.buckets
            run(B, pos_weight, self.dropout_enabled, 1)
            grad_hook.start_bucket(eng, lo1, hi1)
            run(B, pos_weight, self.dropout_enabled, 2)
            grad_hook.finish(eng, lo2, hi2)
        else:
            run(B, pos_weight, self.dropout_enabled)
        if run == eng.train_body:
            self._eager_steps += 1
        loss, logits = eng.loss, eng.logits_view(B)
        self._attach_grads()
        if grad_hook is not None and not overlap:
            grad_hook(eng)
        optimizer.fused_step(eng, advance_rng=True)     # one launch advances the Adam step and the Philox step
        return loss, logits


class THAT_COUNT_PRED(THAT):
    """model/that_count_pred.py:180-302: the count-regression sibling.  Layer for layer the same network as THAT (same
    ``state_dict`` keys and constructor RNG order); ``var_y_shape[-1]`` is the number of activities (9) and the model is
    trained with ``SmoothL1Loss`` on per-activity head counts (``var_mode="count_classification"``), so every backbone
    kernel is reused unchanged and only the loss kernel differs (``csi_smooth_l1``)."""


class THAT_MULTI_HEAD(THAT):
    """model/that_multi_head.py:180-306: the five-head sibling.  Same backbone as THAT; ``layer_output`` is a ModuleList of
    five ``Linear(288, out)`` heads, registered (and initialised) BEFORE the backbone, so ``state_dict`` has the reference's
    171 keys in the reference's order.  ``forward`` returns ``[B, 5, out]`` (that_multi_head.py:304-305).  It is trained
    with ``PermutationMatchingLoss``; the five heads run as one GEMM over a padded [5*cp, 288] operand and the loss /
    its gradient are one kernel (``csi_perm_ce``)."""
    NUM_OUTPUT_HEADS = 5


class _PermCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        from .ops import NativeOps
        B, H, C = pred.shape
        cp = LY.ru(C, 16)
        z = torch.zeros(B, H * cp, device=pred.device)
        z.view(B, H, cp)[:, :, :C] = pred.detach().float()
        y = target.detach().reshape(B, H * C).float().contiguous()
        loss = torch.zeros(1, device=pred.device)
        dz = torch.zeros(B, H * cp, device=pred.device)
        NativeOps(pred.device).perm_ce(z, y, B, H, C, cp, 1.0, loss, dz)
        ctx.save_for_backward(dz.view(B, H, cp)[:, :, :C])
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


class PermutationMatchingLoss(torch.nn.Module):
    """model/that_multi_head.py:309-342 for predictions / targets ``[B, heads, classes]`` on a CUDA device: per sample
    the permutation of the heads with the smallest mean cross-entropy against ``argmax(targets)`` (first minimum in
    ``itertools.permutations`` order), loss = mean cross-entropy of the matched heads.  One kernel (``csi_perm_ce``)
    instead of the reference's B x 120 Python loop; ``train()`` / ``fused_train_step(loss_kind="perm_ce")`` fuse it."""

    def forward(self, predictions, targets):
        if predictions.dim() != 3 or predictions.shape != targets.shape:
            raise ValueError(f"expected predictions and targets [B, heads, classes], got {tuple(predictions.shape)} "
                             f"and {tuple(targets.shape)}")
        if predictions.device.type != "cuda":
            raise RuntimeError("PermutationMatchingLoss runs csi_perm_ce on a CUDA device; there is no CPU path")
        return _PermCE.apply(predictions, targets)
