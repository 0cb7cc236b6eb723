"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference CNN_2D (CSI-as-image) path, SURVEY.md section 8(f)-3.

The oracle of BASELINE config 4 (product: multi_modal_csi_b200/cnn2d.py + csrc/cnn2d.cu).  Nothing in the product package
imports this file.

Restates, as pure functions over a ``state_dict`` (plain torch CPU ops):

  * benchmark/wifi_csi/model/cnn_2d.py:23-66   CNN_2D.__init__ (layer shapes, initialisation order) -> :func:`cnn2d_init`
  * benchmark/wifi_csi/model/cnn_2d.py:70-99   CNN_2D.forward                                       -> :func:`cnn2d_forward`
  * benchmark/wifi_csi/model/cnn_2d.py:162-166 Adam(weight_decay=1e-4), BCEWithLogitsLoss(pos_weight=6)

Pinned by ``tests/golden/cnn2d_anchor.npz`` (``oracle/make_golden.py::cnn2d_case``, generated from the unmodified
reference): per-tensor checksums of the initial weights under the reference seed, logits, loss and gradient norms.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as Fn

LEAKY = 0.01        # torch.nn.LeakyReLU() default, cnn_2d.py:59
P_DROP = 0.2        # cnn_2d.py:61
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
CONVS = ((1, 32, 27, 7), (32, 64, 15, 3), (64, 128, 7, 1))       # (in, out, kernel, stride), cnn_2d.py:42-55


def _default_conv_init(shape):
    """torch.nn.Conv2d / Linear.reset_parameters(): kaiming_uniform_(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias."""
    w = torch.empty(*shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    b = torch.empty(shape[0]).uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
    return w, b


def cnn2d_init(out: int) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of a freshly constructed CNN_2D, drawing from the global CPU RNG in the reference's order:
    the four BatchNorm2d (no draws), the three Conv2d and the Linear with their default initialisation, then
    xavier_uniform_ over the four weights (cnn_2d.py:63-66).  Keys follow the reference's registration order."""
    sd = OrderedDict()
    for i, c in enumerate((1, 32, 64, 128)):
        sd[f"layer_norm_{i}.weight"] = torch.ones(c)
        sd[f"layer_norm_{i}.bias"] = torch.zeros(c)
        sd[f"layer_norm_{i}.running_mean"] = torch.zeros(c)
        sd[f"layer_norm_{i}.running_var"] = torch.ones(c)
        sd[f"layer_norm_{i}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for i, (ci, co, k, _s) in enumerate(CONVS):
        sd[f"layer_cnn_2d_{i}.weight"], sd[f"layer_cnn_2d_{i}.bias"] = _default_conv_init((co, ci, k, k))
    sd["layer_linear.weight"], sd["layer_linear.bias"] = _default_conv_init((out, 128))
    for name in ("layer_cnn_2d_0.weight", "layer_cnn_2d_1.weight", "layer_cnn_2d_2.weight", "layer_linear.weight"):
        torch.nn.init.xavier_uniform_(sd[name])
    return sd


def _bn2d(sd, prefix, x, training, update_stats):
    if not training:
        return Fn.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                             sd[prefix + "bias"], False, BN_MOMENTUM, BN_EPS)
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if update_stats:
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            sd[prefix + "running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            sd[prefix + "running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
            sd[prefix + "num_batches_tracked"] += 1
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    return xh * sd[prefix + "weight"][None, :, None, None] + sd[prefix + "bias"][None, :, None, None]


class _RoundBF16(torch.autograd.Function):
    """x -> bf16(x) with a straight-through gradient (optionally rounded to bf16 as well): restates WHERE the bf16 mode of
    the CUDA path stores a tensor in bf16, so that "bf16 parity" can be stated against the reference ALGORITHM evaluated
    with bf16-rounded operands instead of only against a loose tolerance."""

    @staticmethod
    def forward(ctx, x, round_grad):
        ctx.round_grad = round_grad
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().to(g.dtype) if ctx.round_grad else g), None


def cnn2d_forward(sd, x: torch.Tensor, training: bool = False, update_stats: bool = True, drop=None,
                  emulate_bf16: bool = False) -> torch.Tensor:
    """cnn_2d.py:70-99.  x: [B, T, F] -> logits [B, out]; BatchNorm2d -> Conv2d -> LeakyReLU -> Dropout(0.2), three
    times, then BatchNorm2d, mean over the image and the Linear layer.

    emulate_bf16: round to bf16 wherever the bf16 mode of the CUDA path stores bf16 -- the normalised conv inputs (patch
    matrix), the weights, the conv outputs, the activations, the pooled features -- and, in backward, the gradients of
    the conv outputs and of the patch matrices.  Arithmetic stays fp32 (the tensor cores accumulate in fp32)."""
    drop = drop or (lambda t, p: t)
    r = (lambda t, g=False: _RoundBF16.apply(t, g)) if emulate_bf16 else (lambda t, g=False: t)
    t = x.unsqueeze(1)
    for i, (_ci, _co, _k, s) in enumerate(CONVS):
        t = r(_bn2d(sd, f"layer_norm_{i}.", t, training, update_stats), i > 0)
        t = r(Fn.conv2d(t, r(sd[f"layer_cnn_2d_{i}.weight"]), sd[f"layer_cnn_2d_{i}.bias"], stride=s), True)
        t = r(drop(Fn.leaky_relu(t, LEAKY), P_DROP))
    t = _bn2d(sd, "layer_norm_3.", t, training, update_stats)
    return Fn.linear(r(t.mean(dim=(-2, -1))), r(sd["layer_linear.weight"]), sd["layer_linear.bias"])


def bce_with_logits(z, y, pos_weight: float = 6.0):
    """BCEWithLogitsLoss(pos_weight=6) of cnn_2d.py:166."""
    return -(pos_weight * y * Fn.logsigmoid(z) + (1 - y) * Fn.logsigmoid(-z)).mean()


def loss_and_grads(sd, x, y, pos_weight: float = 6.0, emulate_bf16: bool = False):
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    work = OrderedDict(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work.update(leaves)
    logits = cnn2d_forward(work, x, training=True, update_stats=False, emulate_bf16=emulate_bf16)
    loss = bce_with_logits(logits, y, pos_weight)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads))
