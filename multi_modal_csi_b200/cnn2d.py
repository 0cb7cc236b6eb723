"""Drop-in for benchmark/wifi_csi/model/cnn_2d.py (BASELINE config 4, the CSI-as-image path): ``CNN_2D(var_x_shape,
var_y_shape)`` and ``run_cnn_2d(...)``.

Same ``state_dict`` keys, registration order and initialisation draws as the reference (cnn_2d.py:23-66: four
BatchNorm2d, three Conv2d with default init, Linear, then ``xavier_uniform_`` over the four weights), so
``torch.random.manual_seed(r + 39)`` gives the reference's initial weights.  ``forward`` (cnn_2d.py:70-99) and its backward
run on the hand-written sm_100a kernels of ``libcsi_that.so``: the image is kept NHWC, every strided Conv2d is a BatchNorm-
fused im2col + the tcgen05 GEMM (``csrc/cnn2d.cu`` describes the data layout), the loss / Adam / data-parallel layers are
the ones of the THAT path.  No PyTorch-operator or CPU fallback.
"""
from __future__ import annotations

import time
import weakref
from collections import OrderedDict
from typing import Optional

import torch

from . import layout as LY
from .arena_module import ArenaModule
from .engine import StepCounters
from .that import _kaiming_conv

CONVS = ((1, 32, 27, 7), (32, 64, 15, 3), (64, 128, 7, 1))       # (in, out, kernel, stride), cnn_2d.py:42-55
P_DROP = 0.2                                                     # cnn_2d.py:61
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
SITE0 = 9100                                                     # Philox stream ids of the three dropout layers


class Geom2D:
    """Image sizes through the three valid, strided convolutions."""

    def __init__(self, T: int, F: int, out: int):
        self.T, self.F, self.out = T, F, out
        self.H, self.W, self.C = [T], [F], [1]
        for (ci, co, k, s) in CONVS:
            if self.H[-1] < k or self.W[-1] < k:
                raise ValueError(f"input {T}x{F} is too small for the {k}x{k} convolution (cnn_2d.py:42-55)")
            self.H.append((self.H[-1] - k) // s + 1)
            self.W.append((self.W[-1] - k) // s + 1)
            self.C.append(co)
        self.Kp = [LY.ru(k * k * ci, 16) for (ci, co, k, s) in CONVS]
        self.ld_out = LY.ru(out, 16)

    def rows(self, i: int, B: int) -> int:
        """rows of the NHWC matrix that ENTERS conv i (i = 3: the final feature map)"""
        return B * self.H[i] * self.W[i]


def parameter_specs(out: int):
    specs, bufs = OrderedDict(), OrderedDict()
    for i, c in enumerate((1, 32, 64, 128)):
        specs[f"layer_norm_{i}.weight"] = (c,)
        specs[f"layer_norm_{i}.bias"] = (c,)
        bufs[f"layer_norm_{i}.running_mean"] = torch.zeros(c)
        bufs[f"layer_norm_{i}.running_var"] = torch.ones(c)
        bufs[f"layer_norm_{i}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for i, (ci, co, k, _s) in enumerate(CONVS):
        specs[f"layer_cnn_2d_{i}.weight"] = (co, ci, k, k)
        specs[f"layer_cnn_2d_{i}.bias"] = (co,)
    specs["layer_linear.weight"] = (out, 128)
    specs["layer_linear.bias"] = (out,)
    return specs, bufs


def _initial_values(out: int):
    """Draws from the global CPU RNG in the order of the reference constructor (cnn_2d.py:37-66)."""
    vals = OrderedDict()
    for i, c in enumerate((1, 32, 64, 128)):
        vals[f"layer_norm_{i}.weight"] = torch.ones(c)
        vals[f"layer_norm_{i}.bias"] = torch.zeros(c)
    for i, (ci, co, k, _s) in enumerate(CONVS):
        vals[f"layer_cnn_2d_{i}.weight"], vals[f"layer_cnn_2d_{i}.bias"] = _kaiming_conv((co, ci, k, k))
    vals["layer_linear.weight"], vals["layer_linear.bias"] = _kaiming_conv((out, 128))
    for name in ("layer_cnn_2d_0.weight", "layer_cnn_2d_1.weight", "layer_cnn_2d_2.weight", "layer_linear.weight"):
        torch.nn.init.xavier_uniform_(vals[name])
    return vals


class CNN2DEngine(StepCounters):
    """Buffers + launch sequence of one CNN_2D train / eval step on one GPU (the counterpart of engine.THATEngine)."""

    def __init__(self, geom: Geom2D, max_batch: int, params, grads, arena, buffers, act_dtype, ops=None, rng=None, opt_step=None):
        self.g, self.B = geom, int(max_batch)
        self.params, self.grads, self.arena, self.bn = params, grads, arena, buffers
        self.dev, self.adt = params.device, act_dtype
        if ops is None:
            from .ops import NativeOps                   # raises without libcsi_that.so or on a non-CUDA device
            ops = NativeOps(self.dev)
        self.ops = ops
        self.rng = rng if rng is not None else torch.tensor([0, 0], dtype=torch.int64, device=self.dev)
        self.opt_step = opt_step if opt_step is not None else torch.ones(1, dtype=torch.int64, device=self.dev)
        self.rng_used = False
        self.weights_dirty = True
        self.loss_kind = "bce"
        self._graphs, self._graph_launches = {}, {}
        self._alloc()

    def P(self, name):
        off, shp = self.arena.offsets[name], self.arena.shapes[name]
        return self.params[off:off + LY.numel(shp)]

    def G(self, name):
        off, shp = self.arena.offsets[name], self.arena.shapes[name]
        return self.grads[off:off + LY.numel(shp)]

    def _alloc(self):
        g, B, dev, adt, f32 = self.g, self.B, self.dev, self.adt, torch.float32
        z = lambda *shape, dtype=f32: torch.zeros(*shape, dtype=dtype, device=dev)
        nstat = sum(2 * c for c in g.C)
        self.stat_pool = z(nstat, dtype=torch.float64)        # forward BatchNorm sums of the four layers
        self.red_pool = z(nstat + 2 * 32, dtype=torch.float64)     # backward sums (+ the conv-0 column sums)
        self.L = []
        so = 0
        for i in range(4):
            C = g.C[i]
            layer = {"sums": self.stat_pool[so:so + 2 * C], "red": self.red_pool[so:so + 2 * C],
                     "mean": z(C), "invstd": z(C), "scale": z(C), "shift": z(C)}
            so += 2 * C
            if i < 3:
                ci, co, k, s = CONVS[i]
                M = g.rows(i + 1, B)
                layer.update({
                    "col": z(M, g.Kp[i], dtype=adt), "z": z(M, co, dtype=adt), "y": z(M, co, dtype=adt),
                    "mask": z(M * co // 8, dtype=torch.uint8), "gz": z(M, co, dtype=adt),
                    "wf": z(co, g.Kp[i], dtype=adt), "wscratch": z(co, g.Kp[i]),
                })
                if i > 0:
                    layer.update({"wb": z(g.Kp[i], co, dtype=adt), "gcol": z(M, g.Kp[i], dtype=adt),
                                  "gt": z(g.rows(i, B), ci)})
            self.L.append(layer)
        self.red0 = self.red_pool[nstat:nstat + 64]
        self.zeros32, self.ones32 = z(32), torch.ones(32, device=dev)
        self.feat, self.featd = z(B, 128), z(B, 128, dtype=adt)
        self.dfeat = z(B, 128)
        self.wf_lin, self.wb_lin = z(g.out, 128, dtype=adt), z(128, g.ld_out, dtype=adt)
        self.logits, self.dlogits = z(B, g.ld_out), z(B, g.ld_out)
        self.dlogits_a = z(B, g.ld_out, dtype=adt)
        self.loss = z(1)
        self.y_static = z(B, g.out)
        self.x_static = None                                  # dense batch [B,T,F]: allocated on the first gather / augmentation

    # ------------------------------------------------------------------ weights
    def repack(self):
        ops, g = self.ops, self.g
        for i, (ci, co, k, _s) in enumerate(CONVS):
            L = self.L[i]
            ops.conv2d_pack(self.P(f"layer_cnn_2d_{i}.weight"), co, ci, k, L["wf"], g.Kp[i], L.get("wb"), co)
        ops.conv2d_pack(self.P("layer_linear.weight"), g.out, 128, 1, self.wf_lin, 128, self.wb_lin, g.ld_out)
        self.weights_dirty = False

    def _bn(self, i, x, rows, training):
        """BatchNorm2d i (cnn_2d.py:75,80,85,90): batch statistics of x [rows, C_i] -> mean / invstd / fused scale, shift."""
        ops, L, C = self.ops, self.L[i], self.g.C[i]
        p = f"layer_norm_{i}."
        if training:
            ops.nhwc_stats(x, rows, C, L["sums"])
        ops.bn2d_finalize(L["sums"], C, rows, self.P(p + "weight"), self.P(p + "bias"), self.bn[p + "running_mean"],
                          self.bn[p + "running_var"], self.bn[p + "num_batches_tracked"], BN_MOMENTUM, BN_EPS, training,
                          L["mean"], L["invstd"], L["scale"], L["shift"])

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, B: int, training: bool, dropout: bool = True) -> torch.Tensor:
        """x: fp32 [B, T, F] on the device -> logits [B, out] (a view of the engine's static buffer)."""
        ops, g = self.ops, self.g
        assert B <= self.B
        if self.weights_dirty:
            self.repack()
        if training:
            ops.fill_f64(self.stat_pool)
        pd = P_DROP if (training and dropout) else 0.0
        inp = x
        self._bn(0, x, g.rows(0, B), training)
        for i, (ci, co, k, s) in enumerate(CONVS):
            L = self.L[i]
            M = g.rows(i + 1, B)
            ops.im2col_bn(inp, B, g.H[i], g.W[i], ci, k, s, L["scale"], L["shift"], L["col"], g.Kp[i])
            ops.alg_flops = 2 * M * co * k * k * ci
            ops.gemm_nt(L["col"], L["wf"], L["z"], M, co, [(0, 0, 0, g.Kp[i])], self.P(f"layer_cnn_2d_{i}.bias"), None, 0.0, 0,
                        self.rng)
            ops.act_drop_fwd(L["z"], L["y"], M * co, pd, SITE0 + i, self.rng, L["mask"] if pd > 0 else None)
            self._bn(i + 1, L["y"], M, training)
            inp = L["y"]
        P = g.H[3] * g.W[3]
        ops.pool_bn_fwd(inp, B, P, 128, self.L[3]["scale"], self.L[3]["shift"], self.feat, self.featd)
        ops.alg_flops = 2 * B * 128 * g.out
        ops.gemm_nt(self.featd, self.wf_lin, self.logits, B, g.out, [(0, 0, 0, 128)], self.P("layer_linear.bias"), None, 0.0, 0,
                    self.rng)
        return self.logits[:B, :g.out]

    def logits_view(self, B):
        return self.logits[:B, :self.g.out]

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits: Optional[torch.Tensor], B: int, dropout: bool = True, zero_grads: bool = True):
        """Gradients of every parameter into the flat arena from dL/dlogits (None = the buffer written by loss_fwd_bwd)."""
        ops, g = self.ops, self.g
        pd = P_DROP if dropout else 0.0
        if zero_grads:
            ops.fill_f32(self.grads, 0.0)
        ops.fill_f64(self.red_pool)
        if dlogits is not None:
            self.dlogits[:B, :g.out].copy_(dlogits)
        ops.dropout_rows(self.dlogits, self.dlogits_a, B, g.ld_out, 0.0, 0, self.rng)            # cast to the act dtype
        ops.alg_flops = 2 * B * 128 * g.out
        ops.gemm_tn(self.dlogits_a, self.featd, self.G("layer_linear.weight"), 128, 1, B, g.out, [(0, 0, 0, 128)])
        ops.colsum_tokens(self.dlogits_a, B, 1, 0, g.out, self.G("layer_linear.bias"))
        ops.alg_flops = 2 * B * 128 * g.out
        ops.gemm_nt(self.dlogits_a, self.wb_lin, self.dfeat, B, 128, [(0, 0, 0, g.ld_out)], None, None, 0.0, 0, self.rng)
        # final BatchNorm2d + spatial mean: the gradient of feat is shared by the P positions (scaled 1/P)
        P = g.H[3] * g.W[3]
        gup, g_div, g_scale = self.dfeat, P, 1.0 / P
        for i in (2, 1, 0):
            ci, co, k, s = CONVS[i]
            L, Ln = self.L[i], self.L[i + 1]
            M = g.rows(i + 1, B)
            pn = f"layer_norm_{i + 1}."
            # BatchNorm2d i+1 backward fused with the Dropout + LeakyReLU backward of block i -> gz_i
            ops.bn2d_bwd_reduce(gup, g_div, g_scale, L["y"], M, co, Ln["mean"], Ln["invstd"], Ln["red"])
            ops.bn2d_bwd_apply(gup, g_div, g_scale, L["y"], L["z"], L["mask"] if pd > 0 else None, pd, M, co, Ln["mean"],
                               Ln["invstd"], self.P(pn + "weight"), Ln["red"], L["gz"], self.G(pn + "weight"), self.G(pn + "bias"))
            # weight gradient: contraction over the M patch rows, then back to the reference [N, C, kh, kw] layout
            ops.fill_f32(L["wscratch"], 0.0)
            ops.alg_flops = 2 * M * co * k * k * ci
            ops.gemm_tn(L["gz"], L["col"], L["wscratch"], g.Kp[i], 1, M, co, [(0, 0, 0, g.Kp[i])])
            ops.conv2d_unpack_grad(L["wscratch"], co, ci, k, g.Kp[i], self.G(f"layer_cnn_2d_{i}.weight"))
            if i > 0:
                ops.colsum_tokens(L["gz"], 1, M, 0, co, self.G(f"layer_cnn_2d_{i}.bias"))
                ops.alg_flops = 2 * M * co * k * k * ci
                ops.gemm_nt(L["gz"], L["wb"], L["gcol"], M, g.Kp[i], [(0, 0, 0, co)], None, None, 0.0, 0, self.rng)
                ops.col2im(L["gcol"], B, g.H[i], g.W[i], ci, k, s, g.Kp[i], L["gt"])
                gup, g_div, g_scale = L["gt"], 1, 1.0
            else:
                # single-channel BatchNorm2d 0 + conv-0 bias from the column sums of gz_0 and gz_0 * z_0 (cnn2d.cu)
                ops.bn2d_bwd_reduce(L["gz"], 1, 1.0, L["z"], M, co, self.zeros32, self.ones32, self.red0)
                ops.bn0_grads(self.red0, L["wf"], g.Kp[0], co, k * k, self.P("layer_cnn_2d_0.bias"), self.P("layer_norm_0.weight"),
                              self.P("layer_norm_0.bias"), self.G("layer_norm_0.weight"), self.G("layer_norm_0.bias"),
                              self.G("layer_cnn_2d_0.bias"))

    # ------------------------------------------------------------------ loss / fused step body
    def loss_fwd_bwd(self, y, B, pos_weight=6.0, grad_scale=1.0, want_grad=True):
        self.ops.bce_logits(self.logits, y, B, self.g.out, pos_weight, grad_scale, self.loss, self.dlogits if want_grad else None)
        return self.loss

    def train_body(self, x, B, pos_weight, dropout):
        self.repack()
        self.forward(x, B, True, dropout)
        self.loss_fwd_bwd(self.y_static, B, pos_weight)
        self.backward(None, B, dropout=dropout, zero_grads=True)


class _CNN2DFunction(torch.autograd.Function):
    """Autograd bridge for the reference-style loop ``loss(model(x), y).backward()`` (train.py:96-100)."""

    @staticmethod
    def forward(ctx, x, anchor, model):
        B = x.shape[0]
        eng = model._engine_for(B)
        eng.repack()
        eng.begin_train_forward()
        logits = eng.forward(x, B, training=True, dropout=model.dropout_enabled)
        ctx.model, ctx.B = model, B
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        model, B = ctx.model, ctx.B
        eng = model._engine
        fresh = all(p.grad is None for p in model.parameters())
        eng.backward(dlogits.contiguous().float(), B, dropout=model.dropout_enabled, zero_grads=fresh)
        eng.end_train_step()
        model._attach_grads()
        return None, None, None


class CNN_2D(ArenaModule):
    """``CNN_2D(var_x_shape, var_y_shape)``: var_x_shape[-2:] = (T, F), var_y_shape[-1] = out (cnn_2d.py:26-36)."""

    def __init__(self, var_x_shape, var_y_shape, act_dtype: Optional[str] = None, max_batch: Optional[int] = None):
        super().__init__()
        T, F, out = int(var_x_shape[-2]), int(var_x_shape[-1]), int(var_y_shape[-1])
        self.geom = Geom2D(T, F, out)
        self.specs, bufs = parameter_specs(out)
        self.arena = LY.build_arena(self.specs)
        self.act_dtype = {None: torch.bfloat16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
                          "fp32": torch.float32, "float32": torch.float32}[act_dtype]
        self.max_batch = max_batch
        self.dropout_enabled = True
        self._rng_seed = int(torch.initial_seed() & 0x7FFFFFFF)
        self._rng = None
        self._opt_step = None
        self._engine = None
        self._ops_override = None
        vals = _initial_values(out)
        flat = torch.zeros(self.arena.size)
        object.__setattr__(self, "_flat", flat)
        object.__setattr__(self, "_gflat", torch.zeros(self.arena.size))
        # registration order of the reference: the four norms (parameters + buffers), the convolutions, the linear layer
        for name, shape in self.specs.items():
            off = self.arena.offsets[name]
            data = flat[off:off + LY.numel(shape)].view(shape)
            data.copy_(vals[name])
            p = torch.nn.Parameter(data)
            p._csi_owner = weakref.ref(self)
            p._csi_name = name
            self._register(name, p, is_buffer=False)
            if name.startswith("layer_norm_") and name.endswith(".bias"):
                pre = name[:-4]
                for leaf in ("running_mean", "running_var", "num_batches_tracked"):
                    self._register(pre + leaf, bufs[pre + leaf], is_buffer=True)

    def configure(self, act_dtype: Optional[str] = None, max_batch: Optional[int] = None):
        if act_dtype is not None:
            self.act_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[act_dtype]
        if max_batch is not None:
            self.max_batch = max_batch
        self._engine = None
        return self

    def _engine_for(self, B: int) -> CNN2DEngine:
        eng = self._engine
        if eng is None or eng.B < B:
            if self._flat.device.type != "cuda" and self._ops_override is None:
                raise RuntimeError("multi_modal_csi_b200.CNN_2D runs only on a CUDA (sm_100a) device: call .to('cuda') first; "
                                   "there is no CPU path")
            mb = max(B, self.max_batch or 0)
            rng, opt_step = self._counters(self._flat.device)
            bn = OrderedDict((k, v) for k, v in self.named_buffers())
            eng = CNN2DEngine(self.geom, mb, self._flat, self._gflat, self.arena, bn, self.act_dtype, ops=self._ops_override,
                              rng=rng, opt_step=opt_step)
            if self._engine is not None:
                eng.rng_used = self._engine.rng_used
            self._engine = eng
        return eng

    def forward(self, var_input: torch.Tensor) -> torch.Tensor:
        """float32 [B, T, F] -> float32 logits [B, out]  (cnn_2d.py:70-99)."""
        x = var_input
        if x.dim() != 3 or x.shape[1] != self.geom.T or x.shape[2] != self.geom.F:
            raise ValueError(f"expected input [B,{self.geom.T},{self.geom.F}], got {tuple(x.shape)}")
        x = x.float().contiguous()
        if x.device != self._flat.device:
            raise RuntimeError(f"input on {x.device} but model on {self._flat.device}")
        if self.training and torch.is_grad_enabled():
            anchor = torch.zeros((), device=x.device, requires_grad=True)
            return _CNN2DFunction.apply(x, anchor, self)
        N = x.shape[0]
        eng = self._engine_for(min(N, self.max_batch or 128))
        eng.repack()
        outs = []
        for i in range(0, N, eng.B):
            xb = x[i:i + eng.B]
            if self.training:
                eng.begin_train_forward()
            outs.append(eng.forward(xb, xb.shape[0], training=self.training, dropout=self.dropout_enabled).clone())
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def fused_train_step(self, x, y, optimizer, pos_weight: float = 6.0, augment: bool = True, grad_hook=None, offs=None,
                         lens=None, loss_kind: str = "bce", **_unused):
        """augmentation + forward + BCEWithLogits(pos_weight) + backward + Adam as one launch sequence (train.py:84-101 with
        the loss and optimizer of cnn_2d.py:162-166).  x: fp32 [B,T,F] on the device, or a packed arena with offs / lens
        (loader.CSIBatchSource).  Returns (loss, logits) views of static buffers."""
        if loss_kind != "bce":
            raise ValueError("CNN_2D is trained with BCEWithLogitsLoss (cnn_2d.py:166)")
        B = y.shape[0]
        eng = self._engine_for(B)
        g = self.geom
        yf = y.reshape(B, -1).float()
        eng.begin_train_forward()
        eng.ops.copy_f32(eng.y_static, yf.contiguous(), B * g.out)
        if offs is not None or augment:
            if eng.x_static is None:
                eng.x_static = torch.zeros(eng.B, g.T, g.F, device=eng.dev)
            eng.ops.gather_aug(x, offs, lens, B, g.T, g.F, eng.x_static, augment, eng.rng)
            x = eng.x_static
        else:
            x = x.float().contiguous()
        eng.train_body(x, B, pos_weight, self.dropout_enabled)
        self._attach_grads()
        if grad_hook is not None:
            grad_hook(eng)
        optimizer.fused_step(eng, advance_rng=True)
        return eng.loss, eng.logits_view(B)


def run_cnn_2d(data_train_x, data_train_y, data_test_x, data_test_y, var_repeat=10):
    """cnn_2d.py:103-230: same preprocessing, seeds, optimizer (Adam, weight_decay 1e-4 -> FusedAdam), loss
    (BCEWithLogits, pos_weight 6), train() call and result dictionary."""
    from sklearn.metrics import accuracy_score, classification_report
    from torch.utils.data import TensorDataset
    from .optim import FusedAdam
    from .preset import preset
    from .train import train
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    data_train_x = data_train_x.reshape(data_train_x.shape[0], data_train_x.shape[1], -1)
    data_test_x = data_test_x.reshape(data_test_x.shape[0], data_test_x.shape[1], -1)
    var_x_shape, var_y_shape = data_train_x[0].shape, data_train_y[0].reshape(-1).shape
    data_train_set = TensorDataset(torch.from_numpy(data_train_x), torch.from_numpy(data_train_y))
    data_test_set = TensorDataset(torch.from_numpy(data_test_x), torch.from_numpy(data_test_y))
    result = {}
    result_accuracy, result_time_train, result_time_test = [], [], []
    for var_r in range(var_repeat):
        print("Repeat", var_r)
        torch.random.manual_seed(var_r + 39)
        model_cnn_2d = CNN_2D(var_x_shape, var_y_shape, act_dtype=preset["nn"].get("dtype", "bf16"),
                              max_batch=preset["nn"]["batch_size"]).to(device)
        optimizer = FusedAdam(model_cnn_2d.parameters(), lr=preset["nn"]["lr"], weight_decay=1e-4)
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([6] * var_y_shape[-1]).to(device))
        var_time_0 = time.time()
        var_best_weight = train(model=model_cnn_2d, optimizer=optimizer, loss=loss, data_train_set=data_train_set,
                                data_test_set=data_test_set, var_threshold=preset["nn"]["threshold"],
                                var_batch_size=preset["nn"]["batch_size"], var_epochs=preset["nn"]["epoch"], device=device,
                                var_mode="baseline")
        var_time_1 = time.time()
        model_cnn_2d.load_state_dict(var_best_weight)
        model_cnn_2d.eval()
        with torch.no_grad():
            predict_test_y = model_cnn_2d(torch.from_numpy(data_test_x).to(device))
        predict_test_y = (torch.sigmoid(predict_test_y) > preset["nn"]["threshold"]).float().cpu().numpy()
        var_time_2 = time.time()
        data_test_y_c = data_test_y.reshape(-1, data_test_y.shape[-1])
        predict_test_y_c = predict_test_y.reshape(-1, data_test_y.shape[-1])
        result_acc = accuracy_score(data_test_y_c.astype(int), predict_test_y_c.astype(int))
        result_dict = classification_report(data_test_y_c, predict_test_y_c, digits=6, zero_division=0, output_dict=True)
        result["repeat_" + str(var_r)] = result_dict
        result_accuracy.append(result_acc)
        result_time_train.append(var_time_1 - var_time_0)
        result_time_test.append(var_time_2 - var_time_1)
        print("repeat_" + str(var_r), result_accuracy)
    result["accuracy"] = {"avg": sum(result_accuracy) / len(result_accuracy), "std": float(torch.tensor(result_accuracy).std()) if len(result_accuracy) > 1 else 0.0}
    result["time_train"] = {"avg": sum(result_time_train) / len(result_time_train)}
    result["time_test"] = {"avg": sum(result_time_test) / len(result_time_test)}
    return result
