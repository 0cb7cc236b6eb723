"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference THAT train step.

This file is the parity oracle for the CUDA path.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs;
the product package (``multi_modal_csi_b200``) never imports it.

It restates, as pure functions over a ``state_dict`` (plain ``torch`` CPU ops, fp32 or fp64),
the algorithm of the reference files below (paths relative to /root/reference):

  * benchmark/wifi_csi/model/that.py:31-90    Gaussian_Position   -> :func:`gaussian_position`
  * benchmark/wifi_csi/model/that.py:100-170  Encoder             -> :func:`encoder`
  * benchmark/wifi_csi/model/that.py:180-302  THAT.forward        -> :func:`that_forward`
  * benchmark/wifi_csi/model/that.py:401      BCEWithLogitsLoss(pos_weight=4) -> :func:`bce_with_logits`
  * benchmark/wifi_csi/train.py:65-73         apply_augmentation  -> :func:`apply_augmentation`
  * benchmark/wifi_csi/train.py:84-101        one train step      -> :func:`train_step`
  * benchmark/wifi_csi/that.py:393-397        Adam(lr, weight_decay) coupled L2 -> :func:`adam_update`
  * benchmark/wifi_csi/utils.py:147-183,213-270  prediction rule + metrics -> :func:`predict_counts`
  * benchmark/wifi_csi/load_data.py:62-78     front zero-pad to T -> :func:`front_pad`
  * benchmark/wifi_csi/model/that_multi_head.py:194-196,304-342  five output heads + PermutationMatchingLoss
                                              -> :func:`that_forward` (stacked heads), :func:`permutation_matching_loss`

Pinning: the reference ships no tests or golden vectors for this path, so the
restatement is pinned against (1) the unmodified reference executed in the build container
(``tests/test_oracle_vs_reference.py``, skipped where /root/reference is absent) and (2) the committed
fixtures in ``tests/golden/`` that ``oracle/make_golden.py`` generated from the unmodified reference.
The underlying arithmetic lives in PyTorch (third-party; reference pins pytorch 2.0.1, this image
has 2.11.0): the oracle is "reference algorithm on torch CPU".
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as Fn

LEFT_KERNELS = (1, 3, 5)      # that.py:200-202
RIGHT_KERNELS = (1, 2, 3)     # that.py:224-226
NUM_LEFT = 4                  # that.py:198
NUM_RIGHT = 1                 # that.py:222
NUM_HEAD = 10                 # that.py:201,225
POOL = 20                     # that.py:196,220
LN_EPS = 1e-6                 # that.py:112,120,206,229
BN_EPS = 1e-5                 # torch default, that.py:130
BN_MOMENTUM = 0.1
LEAKY = 0.01                  # torch default negative_slope, that.py:132,242


# ----------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------
def gaussian_position(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """that.py:61-90 -- softmax over K of the Gaussian log-density, times the embedding, broadcast-added."""
    pos = sd[prefix + "var_position"]            # [L, K] (frozen)
    mu = sd[prefix + "var_mu"]                   # [1, K]
    sigma = sd[prefix + "var_sigma"]             # [1, K]
    emb = sd[prefix + "var_embedding"]           # [K, F]
    diff = pos - mu
    logp = -(diff * diff) / sigma / sigma / 2 - torch.log(sigma)      # that.py:67-73 (same op order)
    w = torch.softmax(logp, dim=-1)
    return x + (w @ emb).unsqueeze(0)


def _mha(sd, prefix, x, num_head=NUM_HEAD):
    """nn.MultiheadAttention(d, 10, batch_first=True), self-attention, dropout 0 (that.py:113-115,149)."""
    B, L, d = x.shape
    hd = d // num_head
    qkv = Fn.linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"])
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):
        return t.reshape(B, L, num_head, hd).transpose(1, 2)          # [B,H,L,hd]

    q, k, v = heads(q), heads(k), heads(v)
    s = (q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, d)
    return Fn.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def _same_conv(x_cf, w, b):
    """Conv1d(padding='same'): left pad (k-1)//2, remainder on the right (k=2 -> 0/1).  x_cf: [B,C,L]."""
    k = w.shape[-1]
    left = (k - 1) // 2
    return Fn.conv1d(Fn.pad(x_cf, (left, k - 1 - left)), w, b)


def _batchnorm(sd, prefix, z, training, update_stats):
    """BatchNorm1d over (B, L) per channel; train mode uses biased batch variance and updates the
    running stats with the unbiased one (torch semantics, that.py:130)."""
    if training:
        mean = z.mean(dim=(0, 2))
        var = z.var(dim=(0, 2), unbiased=False)
        if update_stats:
            n = z.shape[0] * z.shape[2]
            with torch.no_grad():
                sd[prefix + "running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
                sd[prefix + "running_var"].mul_(1 - BN_MOMENTUM).add_(
                    BN_MOMENTUM * var.detach() * (n / max(n - 1, 1)))
                sd[prefix + "num_batches_tracked"].add_(1)
    else:
        mean, var = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    zn = (z - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
    return zn * sd[prefix + "weight"][None, :, None] + sd[prefix + "bias"][None, :, None]


def encoder(sd, prefix, x, kernels, training, update_stats=True, drop=None):
    """that.py:141-170.  ``drop`` is an optional callable(tensor, p) implementing dropout (None = p=0)."""
    drop = drop or (lambda t, p: t)
    d = x.shape[-1]
    t = Fn.layer_norm(x, (d,), sd[prefix + "layer_norm_0.weight"], sd[prefix + "layer_norm_0.bias"], LN_EPS)
    t = _mha(sd, prefix + "layer_attention.", t)
    t = drop(t, 0.1) + x
    s = Fn.layer_norm(t, (d,), sd[prefix + "layer_norm_1.weight"], sd[prefix + "layer_norm_1.bias"], LN_EPS)
    s = s.transpose(1, 2)                                             # [B, d, L]
    acc = 0
    for j, _k in enumerate(kernels):
        cp = f"{prefix}layer_cnn.{j}."
        z = _same_conv(s, sd[cp + "0.weight"], sd[cp + "0.bias"])
        z = _batchnorm(sd, cp + "1.", z, training, update_stats)
        z = Fn.leaky_relu(drop(z, 0.1), LEAKY)                        # Conv -> BN -> Dropout -> LeakyReLU
        acc = acc + z
    s = drop(acc / len(kernels), 0.1).transpose(1, 2)
    return s + t


def _head(sd, name, x_cf):
    z = Fn.conv1d(x_cf, sd[name + ".weight"], sd[name + ".bias"])     # valid conv
    return Fn.leaky_relu(z, LEAKY).sum(dim=-1)


def that_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = False,
                 update_stats: bool = True, drop=None) -> torch.Tensor:
    """that.py:249-302.  x: [B, T, F] -> logits [B, out]."""
    drop = drop or (lambda t, p: t)
    pooled = Fn.avg_pool1d(x.transpose(1, 2), POOL, POOL)             # [B, F, L]   that.py:257-258,279-280
    left = gaussian_position(sd, "layer_left_gaussian.", pooled.transpose(1, 2))
    for i in range(NUM_LEFT):
        left = encoder(sd, f"layer_left_encoder.{i}.", left, LEFT_KERNELS, training, update_stats, drop)
    F_ = left.shape[-1]
    left = Fn.layer_norm(left, (F_,), sd["layer_left_norm.weight"], sd["layer_left_norm.bias"], LN_EPS)
    left = left.transpose(1, 2)
    left = drop(torch.cat([_head(sd, "layer_left_cnn_0", left), _head(sd, "layer_left_cnn_1", left)], -1), 0.5)

    right = pooled
    for i in range(NUM_RIGHT):
        right = encoder(sd, f"layer_right_encoder.{i}.", right, RIGHT_KERNELS, training, update_stats, drop)
    L_ = right.shape[-1]
    right = Fn.layer_norm(right, (L_,), sd["layer_right_norm.weight"], sd["layer_right_norm.bias"], LN_EPS)
    right = right.transpose(1, 2)
    right = drop(torch.cat([_head(sd, "layer_right_cnn_0", right), _head(sd, "layer_right_cnn_1", right)], -1), 0.5)

    feat = torch.cat([left, right], -1)
    if "layer_output.0.weight" in sd:
        # model/that_multi_head.py:194-196,304-305: five Linear(288, out) heads stacked -> [B, 5, out]
        n = 0
        while f"layer_output.{n}.weight" in sd:
            n += 1
        return torch.stack([Fn.linear(feat, sd[f"layer_output.{h}.weight"], sd[f"layer_output.{h}.bias"]) for h in range(n)], 1)
    return Fn.linear(feat, sd["layer_output.weight"], sd["layer_output.bias"])


# ----------------------------------------------------------------------------------------------
# loss / optimizer / augmentation / step
# ----------------------------------------------------------------------------------------------
def bce_with_logits(z: torch.Tensor, y: torch.Tensor, pos_weight: float = 4.0) -> torch.Tensor:
    """mean over B*out of -(w*y*log sigmoid(z) + (1-y)*log sigmoid(-z))   (that.py:401, train.py:97)."""
    return -(pos_weight * y * Fn.logsigmoid(z) + (1 - y) * Fn.logsigmoid(-z)).mean()


def permutation_matching_loss(pred: torch.Tensor, target: torch.Tensor):
    """model/that_multi_head.py:309-342 (PermutationMatchingLoss).  pred, target: [B, H, C]; the target class of slot t is
    argmax(target[b, t]).  Per sample the permutation perm (slot t is scored against head perm[t]) with the smallest mean
    cross-entropy is chosen -- the FIRST one in itertools.permutations order on ties, because the reference keeps the
    incumbent unless ``loss < best`` -- and the loss is the mean cross-entropy of the re-ordered heads over B*H.
    Returns (loss, best permutation per sample [B, H])."""
    from itertools import permutations
    B, H, C = pred.shape
    logp = Fn.log_softmax(pred, dim=-1)                                     # [B, H, C]
    cls = target.argmax(dim=-1)                                             # [B, H]
    # cost[b, h, t] = CE(head h, class of slot t)
    cost = -logp.gather(2, cls.unsqueeze(1).expand(B, H, H))                # [B, H(head), H(slot)]
    perms = torch.tensor(list(permutations(range(H))), dtype=torch.long)    # [P, H]
    slot = torch.arange(H)
    tot = cost.detach()[:, perms, slot].sum(-1) / H                         # [B, P] mean CE of each permutation
    best = perms[tot.argmin(dim=1)]                                         # first minimum = reference tie rule
    loss = cost[torch.arange(B).unsqueeze(1), best, slot].mean()
    return loss, best


def apply_augmentation(x: torch.Tensor, gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """train.py:65-73: +0.1*N(0,1); per-sample scale U[0.9,1.1); Bernoulli(0.96) keep mask (no rescale)."""
    x = x + torch.randn(x.shape, generator=gen, dtype=x.dtype) * 0.1
    scale = torch.rand(x.shape[0], 1, generator=gen, dtype=x.dtype) * 0.2 + 0.9
    x = x * scale.unsqueeze(-1)
    mask = torch.bernoulli(torch.full(x.shape, 0.96, dtype=x.dtype), generator=gen)
    return x * mask


def trainable_names(sd) -> Sequence[str]:
    """Parameter tensors that receive gradients (buffers and the frozen var_position excluded)."""
    skip = ("running_mean", "running_var", "num_batches_tracked", "var_position")
    return [k for k in sd if not k.endswith(skip)]


def adam_update(p, g, m, v, step, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update with coupled L2 (grad += wd * p), in place; step is 1-based."""
    g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def loss_and_grads(sd, x, y, pos_weight=4.0, training=True, update_stats=True):
    """One forward + backward (dropout off).  Returns (logits, loss, {name: grad})."""
    names = trainable_names(sd)
    work = OrderedDict((k, v) for k, v in sd.items())
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    logits = that_forward(work, x, training=training, update_stats=update_stats)
    loss = bce_with_logits(logits, y, pos_weight)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads))


def train_step(sd, opt_state, x, y, lr=5e-4, wd=2e-4, pos_weight=4.0):
    """train.py:96-101 with dropout off and augmentation applied by the caller.  Mutates sd/opt_state."""
    logits, loss, grads = loss_and_grads(sd, x, y, pos_weight, training=True)
    opt_state["step"] = opt_state.get("step", 0) + 1
    for k, g in grads.items():
        if k not in opt_state:
            opt_state[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
        m, v = opt_state[k]
        with torch.no_grad():
            adam_update(sd[k], g, m, v, opt_state["step"], lr, wd)
    return logits, loss


# ----------------------------------------------------------------------------------------------
# loader / prediction rule
# ----------------------------------------------------------------------------------------------
def front_pad(sample: torch.Tensor, T: int) -> torch.Tensor:
    """load_data.py:66-72: zero rows are put in FRONT of a short recording ([t,...] -> [T,...])."""
    t = sample.shape[0]
    if t > T:
        raise ValueError("recording longer than T")                   # np.pad raises on a negative width
    return torch.cat([sample.new_zeros((T - t,) + tuple(sample.shape[1:])), sample], 0)


def predict_counts(logits: torch.Tensor, users: int = 6) -> torch.Tensor:
    """utils.py:234-239,147-183: sigmoid -> per-user argmax, kept if its probability > 0.5 (hard-coded)
    -> per-class counts [N, C]."""
    N = logits.shape[0]
    p = torch.sigmoid(logits.double()).reshape(N, users, -1)
    top, idx = p.max(dim=2)
    onehot = torch.zeros_like(p)
    onehot.scatter_(2, idx.unsqueeze(-1), (top > 0.5).double().unsqueeze(-1))
    return onehot.sum(dim=1)
