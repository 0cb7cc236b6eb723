"""THAT train-step engine: owns the token buffers and composes the C-ABI kernels into
forward / backward / optimizer for one GPU.

Reference path this replaces (benchmark/wifi_csi): model/that.py:249-302 (THAT.forward), autograd of the
same (train.py:100), that.py:401 (BCEWithLogitsLoss), that.py:395-397 + train.py:99-101 (Adam step) and
train.py:65-73 (augmentation).  Every arithmetic step is one call into ``libcsi_that.so`` through
``ops`` (multi_modal_csi_b200/ops.py); this file only sequences them, so a whole train step is a fixed
list of kernel launches that is captured into one CUDA graph (``capture_train_step``).

The ``ops`` object is injectable so that the CPU test-suite can check this sequencing (in particular the
hand-derived backward pass) against the oracle with a torch mirror of the kernels; the product always
uses the native library and fails loudly without it.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import layout as LY
from .layout import GUARD, HALO, ModelGeom, StreamGeom, site

LN_EPS = 1e-6      # that.py:112,120,206,229
BN_EPS = 1e-5      # torch.nn.BatchNorm1d default (that.py:130)
BN_MOMENTUM = 0.1
P_DROP = 0.1       # that.py:117,131,137
P_FEAT = 0.5       # that.py:216,239


class _TokBuf:
    """[GUARD + rows + GUARD, ld] zero-initialised; ``.t`` is the [rows, ld] body."""

    def __init__(self, rows, ld, dtype, device):
        self.full = torch.zeros(rows + 2 * GUARD, ld, dtype=dtype, device=device)
        self.t = self.full[GUARD:GUARD + rows]


class _Range:
    """NVTX range (nsys / ncu --nvtx timelines): per encoder block and pass, only with CSI_NVTX=1 (push/pop cost ~1 us each,
    and they must not be recorded while a CUDA graph is being captured)."""
    ON = os.environ.get("CSI_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Range.ON:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _Range.ON:
            torch.cuda.nvtx.range_pop()


class StepCounters:
    """Philox / Adam step bookkeeping shared by the engines of this package (``rng``, ``opt_step``, ``rng_used``, ``ops``,
    ``params`` and ``grads`` are set by the engine)."""

    # The Philox step is a property of the forward/backward pair, not of the optimizer: backward regenerates the masks
    # its forward drew, then the step moves on -- whichever optimizer (or none) follows.
    def begin_train_forward(self):
        """Call before a train-mode forward: if the previous train forward never reached ``end_train_step`` (no
        backward, e.g. two forwards in a row) its masks must not be reused."""
        if self.rng_used:
            self.ops.advance_counters(self.rng, None)
        self.rng_used = True

    def end_train_step(self):
        """Call after the backward that belongs to the last train forward (autograd path)."""
        if self.rng_used:
            self.ops.advance_counters(self.rng, None)
        self.rng_used = False

    def adam(self, m: torch.Tensor, v: torch.Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
             weight_decay: float = 0.0, grad_scale: float = 1.0, advance_rng: bool = False):
        """One fused Adam launch over the arena; advances the Adam step and, on the fused train step
        (``advance_rng``), the Philox step in the same tiny launch."""
        self.ops.adam_flat(self.params, self.grads, m, v, self.params.numel(), lr, betas[0], betas[1], eps,
                           weight_decay, self.opt_step, grad_scale)
        adv = advance_rng and self.rng_used
        self.ops.advance_counters(self.rng if adv else None, self.opt_step)
        if adv:
            self.rng_used = False
        self.weights_dirty = True
        if advance_rng and hasattr(self, "repack_ahead"):
            self.repack_ahead()                 # fused train step: the operand copies for the next step, off the critical path


class THATEngine(StepCounters):
    def __init__(self, geom: ModelGeom, max_batch: int, params: torch.Tensor, grads: torch.Tensor,
                 arena: LY.Arena, buffers: Dict[str, torch.Tensor], frozen: Dict[str, torch.Tensor],
                 act_dtype: torch.dtype = torch.bfloat16, ops=None, seed: int = 0, rng: Optional[torch.Tensor] = None,
                 opt_step: Optional[torch.Tensor] = None):
        self.g = geom
        self.B = int(max_batch)
        self.params = params              # flat fp32 arena (views of it are the nn.Parameters)
        self.grads = grads                # flat fp32 arena of the same layout
        self.arena = arena
        self.frozen = frozen              # non-trainable parameters (var_position), not in the arena
        self.bn = buffers                 # running_mean / running_var / num_batches_tracked by state_dict key
        self.dev = params.device
        self.adt = act_dtype
        if ops is None:
            from .ops import NativeOps    # raises if libcsi_that.so is missing or the device is not CUDA
            ops = NativeOps(self.dev)
        self.ops = ops
        # the three conv branches of an encoder as one banded GEMM (CSI_NO_FUSED_CONV=1: three launches, A/B runs)
        self.fused_conv = os.environ.get("CSI_NO_FUSED_CONV", "0") != "1"
        self.pack = LY.build_pack_plan(geom, arena, self.fused_conv)
        self.packed = torch.zeros(self.pack.size, dtype=act_dtype, device=self.dev)
        self.pack_table = ops.make_pack_table(self.pack.entries, self.dev)
        self.packed_bias = torch.zeros(max(self.pack.bias_size, 1), dtype=torch.float32, device=self.dev)
        self.bias_table = ops.make_pack_table(self.pack.bias_entries, self.dev)
        self.loss_kind = "bce"                     # "bce" (THAT, that.py:401) | "smooth_l1" (THAT_COUNT_PRED)
        # {seed, step} of the Philox streams and the 1-based Adam step: device tensors owned by the model (THAT._counters)
        # so that a rebuilt engine continues them; a bare engine (tests) makes its own
        self.rng = rng if rng is not None else torch.tensor([seed, 0], dtype=torch.int64, device=self.dev)
        self.opt_step = opt_step if opt_step is not None else torch.ones(1, dtype=torch.int64, device=self.dev)
        self.rng_used = False             # a train-mode forward has drawn masks from the current Philox step
        self._alloc()
        self._graphs = {}
        self._graph_launches = {}
        # left / right streams on two CUDA streams (see _fork); CSI_NO_CONCURRENT=1 for A/B runs
        self.concurrent = os.environ.get("CSI_NO_CONCURRENT", "0") != "1"
        self._side = None
        self._pack_stream = None                   # repack_ahead: packing stream and the not yet awaited pack on it
        self._pack_pending = None
        # A/B switches of the round-2 scheduling changes
        self.pack_ahead_on = os.environ.get("CSI_NO_PACK_AHEAD", "0") != "1"
        self.prefill_on = os.environ.get("CSI_NO_PREFILL", "0") != "1"
        # weight gradients on a third stream (see _wgrad); CSI_NO_WGRAD_STREAM=1 for A/B runs
        self.wgrad_stream_on = os.environ.get("CSI_NO_WGRAD_STREAM", "0") != "1"
        self._wside = None
        self._wgrad_pending = []
        self.weights_dirty = True

    # ------------------------------------------------------------------ views
    def P(self, name):
        if name in self.frozen:
            return self.frozen[name]
        off, shp = self.arena.offsets[name], self.arena.shapes[name]
        return self.params[off:off + LY.numel(shp)]

    def G(self, name):
        off, shp = self.arena.offsets[name], self.arena.shapes[name]
        return self.grads[off:off + LY.numel(shp)]

    def W(self, key):
        m = self.pack.mats[key]
        return self.packed[m.off:m.off + m.rows * m.ld].view(m.rows, m.ld)

    def PB(self, name):
        """Head-padded fp32 copy of a bias vector."""
        m = self.pack.bias_mats[name]
        return self.packed_bias[m.off:m.off + m.rows]

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, B, adt, f32 = self.dev, self.B, self.adt, torch.float32
        self.s: Dict[str, dict] = {}
        # BatchNorm statistics of every encoder live in two fp64 pools (forward sums, backward sums) so that one fill
        # per pass zeroes them all instead of one tiny fill per encoder
        n_stat = sum(sg.n_enc * 2 * 3 * sg.Dp for sg in self.g.streams)
        self.stat_pool = torch.zeros(n_stat, dtype=torch.float64, device=dev)
        self.red_pool = torch.zeros(n_stat, dtype=torch.float64, device=dev)
        stat_off = 0
        for sg in self.g.streams:
            rows, Dp = sg.rows(B), sg.Dp
            st = {"x0": _TokBuf(rows, Dp, f32, dev), "enc": []}
            for _ in range(sg.n_enc):
                st["enc"].append({
                    "t0": _TokBuf(rows, Dp, adt, dev), "mean0": torch.zeros(rows, device=dev),
                    "rstd0": torch.zeros(rows, device=dev),
                    "qkv": _TokBuf(rows, sg.ld3, adt, dev), "o": _TokBuf(rows, sg.dh, adt, dev),
                    "lse": torch.zeros(B * sg.H * sg.L, device=dev),
                    "t": _TokBuf(rows, Dp, f32, dev),
                    "s": _TokBuf(rows, Dp, adt, dev), "mean1": torch.zeros(rows, device=dev),
                    "rstd1": torch.zeros(rows, device=dev),
                    "z": _TokBuf(rows, 3 * Dp, adt, dev),
                    "bn_mean": torch.zeros(3 * Dp, device=dev), "bn_invstd": torch.zeros(3 * Dp, device=dev),
                    "bn_sums": self.stat_pool[stat_off + len(st["enc"]) * 6 * Dp:stat_off + (len(st["enc"]) + 1) * 6 * Dp],
                    "red": self.red_pool[stat_off + len(st["enc"]) * 6 * Dp:stat_off + (len(st["enc"]) + 1) * 6 * Dp],
                    "out": _TokBuf(rows, Dp, f32, dev),
                    # dropout keep bits of the BN block (3 branches + output), written by bn_act_fwd for its backward
                    "dmask": torch.zeros(rows * (Dp // 8), dtype=torch.int32, device=dev),
                })
            st["hn"] = _TokBuf(rows, Dp, adt, dev)
            st["meanf"] = torch.zeros(rows, device=dev)
            st["rstdf"] = torch.zeros(rows, device=dev)
            st["p"] = _TokBuf(rows, 2 * sg.head_np, adt, dev)
            # backward scratch (shared by the encoders of the stream)
            st["dout"] = [_TokBuf(rows, Dp, f32, dev), _TokBuf(rows, Dp, f32, dev)]
            # dz / dtm / dqkv are read by the weight-gradient GEMMs, which run on their own CUDA stream behind the
            # data-gradient chain (see _wgrad): one buffer per encoder, so the next encoder's backward never overwrites
            # an operand a weight gradient is still reading (0.65 GB at B=256, F=270)
            st["dz"] = [_TokBuf(rows, 3 * Dp, adt, dev) for _ in range(sg.n_enc)]
            st["ds"] = _TokBuf(rows, Dp, adt, dev)
            st["dt"] = _TokBuf(rows, Dp, f32, dev)
            st["dtm"] = [_TokBuf(rows, Dp, adt, dev) for _ in range(sg.n_enc)]
            st["do"] = _TokBuf(rows, sg.dh, adt, dev)
            st["dqkv"] = [_TokBuf(rows, sg.ld3, adt, dev) for _ in range(sg.n_enc)]
            st["dt0"] = _TokBuf(rows, Dp, adt, dev)
            st["dp"] = _TokBuf(rows, 2 * sg.head_np, adt, dev)
            st["dhn"] = _TokBuf(rows, Dp, adt, dev)
            stat_off += sg.n_enc * 6 * Dp
            self.s[sg.name] = st
        L, F = self.g.left.L, self.g.F
        self.pe = torch.zeros(L, self.g.left.Dp, device=dev)
        self.pe_w = torch.zeros(L, LY.NUM_GAUSS, device=dev)
        self.dpe_ws = torch.zeros(L, self.g.left.Dp, device=dev)
        self.feat = torch.zeros(B, LY.FEAT, device=dev)
        self.featd = torch.zeros(B, LY.FEAT, dtype=adt, device=dev)
        self.dfeatd = torch.zeros(B, LY.FEAT, device=dev)
        self.dfeat = torch.zeros(B, LY.FEAT, device=dev)
        self.logits = torch.zeros(B, self.g.ld_out, device=dev)
        self.dlogits = torch.zeros(B, self.g.ld_out, device=dev)
        self.dlogits_a = torch.zeros(B, self.g.ld_out, dtype=adt, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.y_static = torch.zeros(B, self.g.n_targets, device=dev)

    def activation_bytes(self) -> int:
        n = 0
        for st in self.s.values():
            def walk(o):
                nonlocal n
                if isinstance(o, _TokBuf):
                    n += o.full.numel() * o.full.element_size()
                elif isinstance(o, torch.Tensor):
                    n += o.numel() * o.element_size()
                elif isinstance(o, dict):
                    for v in o.values():
                        walk(v)
                elif isinstance(o, list):
                    for v in o:
                        walk(v)
            walk(st)
        return n

    def _alg(self, flops):
        """Algorithmic FLOPs (true d, L, k: no head / channel / halo padding) of the NEXT contraction call: the roofline
        numerator ops.NativeOps._work reports for it (BASELINE.md section 3 counts the same way)."""
        self.ops.alg_flops = int(flops)

    # ------------------------------------------------------------------ weights
    def _sync_pack(self):
        """The current stream waits for a ``repack_ahead`` still running on the packing stream."""
        ps = self._pack_pending
        if ps is not None:
            torch.cuda.current_stream(self.dev).wait_stream(ps)
            self._pack_pending = None

    def ensure_packed(self):
        """Operand copies are current (and visible to the current stream) on return."""
        self._sync_pack()
        if self.weights_dirty:
            self.repack()

    def repack_ahead(self):
        """Right after the optimizer: repack on a side stream.  The two pack launches (~50 us) then overlap the next step's
        input stage (pool_dual, HBM-bound, outside the step graph) instead of heading the graph's critical path.  Inline
        packing stays the rule when concurrency is off, on the CPU mirror and while per-launch timing is recorded."""
        if not (self.pack_ahead_on and self.concurrent and self.dev.type == "cuda") or getattr(self.ops, "_prof", None) is not None:
            return
        cur = torch.cuda.current_stream(self.dev)
        if self._pack_stream is None:
            self._pack_stream = torch.cuda.Stream(self.dev)
        ps = self._pack_stream
        ps.wait_stream(cur)
        with torch.cuda.stream(ps):
            self.repack()
        self._pack_pending = ps

    def repack(self):
        """fp32 master weights -> GEMM operand copies (forward + data-gradient layouts) in the act dtype."""
        if self._pack_pending is not None and torch.cuda.current_stream(self.dev) != self._pack_pending:
            self._sync_pack()
        self.ops.pack_weights(self.params, self.packed, self.pack_table, len(self.pack.entries),
                              self.pack.max_elems)
        self.ops.pack_weights(self.params, self.packed_bias, self.bias_table, len(self.pack.bias_entries),
                              max(3 * sg.d for sg in self.g.streams))
        self.weights_dirty = False

    def _bn3(self, sg: StreamGeom, e: int, leaf: str, grad=False, buf=False):
        p = sg.prefix(e)
        out = []
        for j in range(len(sg.kernels)):
            key = f"{p}layer_cnn.{j}.{leaf}"
            out.append(self.bn[key] if buf else (self.G(key) if grad else self.P(key)))
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, B: int, training: bool, dropout: bool = True, augment: bool = False,
                offs=None, lens=None) -> torch.Tensor:
        """x: fp32 [B,T,F] (or a packed ragged arena with offs/lens).  Returns logits [B, out] (a view of
        the engine's static logits buffer)."""
        self.ensure_packed()
        self.forward_input(x, B, training, augment, offs, lens)
        return self.forward_body(B, training, dropout)

    def forward_input(self, x, B, training, augment=False, offs=None, lens=None):
        """Input stage: Gaussian range encoding table + (augment, front-pad, pool, both stream layouts).  The only
        part of a step that depends on the address of the input batch, so it stays outside the captured graph."""
        ops, g = self.ops, self.g
        assert B <= self.B
        gp = "layer_left_gaussian."
        ops.gauss_pe_fwd(self.P(gp + "var_position"), self.P(gp + "var_mu"), self.P(gp + "var_sigma"),
                         self.P(gp + "var_embedding"), g.left.L, LY.NUM_GAUSS, g.F, self.pe_w, self.pe)
        ops.pool_dual(x, offs, lens, B, g.T, g.F, self.pe, self.s["left"]["x0"].t, self.s["right"]["x0"].t,
                      HALO, 1 if (training and augment) else 0, self.rng)

    # ------------------------------------------------------------------ two-stream concurrency
    def _fork(self, fn):
        """Issue ``fn()`` on the engine's second CUDA stream, ordered after everything already on the current stream;
        returns the join (call it where the current stream needs the results).

        The left (temporal) and right (channel) streams of THAT share nothing between the pooling kernel and the
        feature concat (that.py:257-299), so their launch sequences run side by side: the tail wave of a persistent
        GEMM and the HBM-bound LayerNorm/BatchNorm kernels of one stream fill SMs the other leaves idle.  The same
        fork/join is recorded into the CUDA graph during capture.  Sequential (fn runs inline) when ``concurrent`` is
        off, on a non-CUDA device (mirror ops in the CPU tests) or while per-launch timing is being recorded."""
        if not (self.concurrent and self.dev.type == "cuda") or getattr(self.ops, "_prof", None) is not None:
            fn()
            return lambda: None
        cur = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
        return lambda: cur.wait_stream(side)

    def _wgrad(self, fn):
        """Issue the weight-gradient work ``fn()`` (wgrad GEMMs, bias column sums) on the engine's weight-gradient stream,
        ordered after everything already on the current stream.

        Nothing in the backward chain depends on a weight gradient -- only the optimizer / the gradient all-reduce do --
        so these launches leave the critical path (per encoder: 6 GEMMs + 2 column sums, ~0.22 ms of 0.73 ms): they fill
        the SMs whenever the data-gradient chain runs a bandwidth-bound kernel or a tail wave.  ``_wgrad_join`` makes the
        current stream wait for all of them.  Inline when concurrency is off / while per-launch timing is recorded."""
        if not (self.wgrad_stream_on and self.concurrent and self.dev.type == "cuda") or getattr(self.ops, "_prof", None) is not None:
            fn()
            return
        cur = torch.cuda.current_stream(self.dev)
        if self._wside is None:
            self._wside = {}
        ws = self._wside.get(cur.cuda_stream)                   # one weight-gradient stream per issuing (left / right) stream
        if ws is None:
            ws = self._wside[cur.cuda_stream] = torch.cuda.Stream(self.dev)
        ws.wait_stream(cur)
        with torch.cuda.stream(ws):
            fn()
        if ws not in self._wgrad_pending:
            self._wgrad_pending.append(ws)

    def _busy_streams(self):
        """Side streams that may still hold work feeding the gradient arena: the right-stream fork and the weight-gradient streams."""
        out = list(self._wgrad_pending)
        if self._side is not None and self.concurrent:
            out.append(self._side)
        return out

    def _wgrad_join(self):
        """The current stream waits for every weight gradient issued so far (end of a backward part)."""
        if not self._wgrad_pending:
            return
        cur = torch.cuda.current_stream(self.dev)
        for ws in self._wgrad_pending:
            cur.wait_stream(ws)
        self._wgrad_pending = []

    def forward_body(self, B: int, training: bool, dropout: bool = True, prefill: bool = False) -> torch.Tensor:
        """``prefill`` (fused train step): the gradient arena and the backward reduction pool are zeroed here, at the head
        of the second CUDA stream, instead of between the loss and the backward on the critical path."""
        ops, g = self.ops, self.g
        pd = P_DROP if (training and dropout) else 0.0
        pf = P_FEAT if (training and dropout) else 0.0
        if training:
            ops.fill_f64(self.stat_pool, 0.0)

        def right():
            if prefill:
                ops.fill_f32(self.grads, 0.0)
                ops.fill_f64(self.red_pool, 0.0)
            self._forward_stream(1, B, training, pd)
        join = self._fork(right)
        self._forward_stream(0, B, training, pd)
        join()
        ops.alg_scale = 1.0
        ops.dropout_rows(self.feat, self.featd, B, LY.FEAT, pf, LY.SITE_FEAT, self.rng)
        self._alg(2 * B * LY.FEAT * g.out * g.heads)
        if g.heads == 1:
            ops.gemm_nt(self.featd, self.W("f:layer_output.weight"), self.logits, B, g.out, [(0, 0, 0, LY.FEAT)],
                        self.P("layer_output.bias"), None, 0.0, 0, self.rng)
        else:       # the `heads` output layers are one GEMM: head h owns rows h*cp.. of the packed weight (pad rows are zero)
            ops.gemm_nt(self.featd, self.W("f:layer_output.weight"), self.logits, B, g.ld_out, [(0, 0, 0, LY.FEAT)],
                        self.PB("layer_output.bias"), None, 0.0, 0, self.rng)
        return self.logits_view(B)

    def logits_view(self, B: int) -> torch.Tensor:
        """[B, out] (THAT) or [B, heads, out] (multi-head sibling) view of the static logits buffer."""
        g = self.g
        if g.heads == 1:
            return self.logits[:B, :g.out]
        return self.logits[:B].view(B, g.heads, g.cp)[:, :, :g.out]

    def _forward_stream(self, si: int, B: int, training: bool, pd: float):
        """Encoders + head convs + time-sum of one stream (0 = left, 1 = right) into its columns of ``feat``."""
        ops, g = self.ops, self.g
        sg = g.streams[si]
        st = self.s[sg.name]
        rows, d, Dp, L = sg.rows(B), sg.d, sg.Dp, sg.L
        ops.alg_scale = (L / sg.Lp) * (d / Dp)      # roofline numerators count valid tokens and true channels only
        x_in = st["x0"]
        one = [(0, 0, 0, Dp)]
        for e in range(sg.n_enc):
          with _Range(f"fwd/{sg.name}/encoder{e}"):
            a, p = st["enc"][e], sg.prefix(e)
            ops.layernorm_fwd(x_in.t, self.P(p + "layer_norm_0.weight"), self.P(p + "layer_norm_0.bias"),
                              a["t0"].t, a["mean0"], a["rstd0"], B, L, d, HALO, LN_EPS)
            self._alg(2 * B * L * d * 3 * d)
            ops.gemm_nt(a["t0"].t, self.W("f:" + p + "layer_attention.in_proj_weight"), a["qkv"].t, rows,
                        sg.ld3, one, self.PB(p + "layer_attention.in_proj_bias"), None, 0.0, 0, self.rng)
            ops.attn_fwd(a["qkv"].t, a["o"].t, a["lse"], B, L, d, sg.H, sg.hp, HALO)
            self._alg(2 * B * L * d * d)
            ops.gemm_nt(a["o"].t, self.W("f:" + p + "layer_attention.out_proj.weight"), a["t"].t, rows, d,
                        [(0, 0, 0, sg.dh)], self.P(p + "layer_attention.out_proj.bias"), x_in.t, pd,
                        site(si, e, LY.SITE_ATTN), self.rng)
            ops.layernorm_fwd(a["t"].t, self.P(p + "layer_norm_1.weight"), self.P(p + "layer_norm_1.bias"),
                              a["s"].t, a["mean1"], a["rstd1"], B, L, d, HALO, LN_EPS)
            if self.fused_conv:
                # the three Conv1d branches as ONE GEMM with N = 3*Dp over the shared input tile (taps a branch does not
                # have are zero blocks of the fused operand, skipped per column tile)
                segs, bands = sg.conv_bands()
                self._alg(2 * B * L * d * d * sum(sg.kernels))
                ops.gemm_nt_banded(a["s"].t, self.W("f:" + p + "layer_cnn"), a["z"].t, rows, len(sg.kernels) * Dp, segs, bands,
                                   None, None, 0.0, 0, self.rng)
            else:
                for j, k in enumerate(sg.kernels):
                    pl = (k - 1) // 2
                    segs = [(t - pl, 0, t * Dp, Dp) for t in range(k)]
                    self._alg(2 * B * L * d * d * k)
                    ops.gemm_nt(a["s"].t, self.W(f"f:{p}layer_cnn.{j}.0.weight"), a["z"].t[:, j * Dp:], rows, d,
                                segs, None, None, 0.0, 0, self.rng)
            cb = self._bn3(sg, e, "0.bias")
            rm = self._bn3(sg, e, "1.running_mean", buf=True)
            rv = self._bn3(sg, e, "1.running_var", buf=True)
            if training:
                ops.bn_stats(a["z"].t, B, L, HALO, 3 * Dp, a["bn_sums"])
                ops.bn_finalize(a["bn_sums"], Dp, d, 3, B * L, cb, rm, rv,
                                self._bn3(sg, e, "1.num_batches_tracked", buf=True), BN_MOMENTUM, BN_EPS,
                                a["bn_mean"], a["bn_invstd"])
            else:
                ops.bn_eval_prepare(Dp, d, 3, cb, rm, rv, BN_EPS, a["bn_mean"], a["bn_invstd"])
            ops.bn_act_fwd(a["z"].t, a["bn_mean"], a["bn_invstd"], self._bn3(sg, e, "1.weight"),
                           self._bn3(sg, e, "1.bias"), a["t"].t, a["out"].t, B, L, d, HALO, 3,
                           pd, site(si, e, LY.SITE_BRANCH), pd, site(si, e, LY.SITE_SUM), self.rng,
                           a["dmask"] if pd > 0.0 else None)
            x_in = a["out"]
        st["x_last"] = x_in
        nm = f"layer_{sg.name}_norm."
        ops.layernorm_fwd(x_in.t, self.P(nm + "weight"), self.P(nm + "bias"), st["hn"].t, st["meanf"],
                          st["rstdf"], B, L, d, HALO, LN_EPS)
        if self.fused_conv:
            # the two head convs (k = 8 / 16 or 2 / 4 over the same normalised stream) as one banded GEMM
            segs, bands = sg.head_bands()
            self._alg(sum(2 * B * (L - k + 1) * d * sg.head_n * k for k in sg.head_k))
            ops.gemm_nt_banded(st["hn"].t, self.W(f"f:layer_{sg.name}_cnn"), st["p"].t, rows, len(sg.head_k) * sg.head_np,
                               segs, bands, self.PB(f"layer_{sg.name}_cnn.bias"), None, 0.0, 0, self.rng)
        else:
          for j, k in enumerate(sg.head_k):
            w = f"layer_{sg.name}_cnn_{j}"
            segs = [(t, 0, t * Dp, Dp) for t in range(k)]
            self._alg(2 * B * (L - k + 1) * d * sg.head_n * k)          # valid convolution: L - k + 1 outputs per sample
            ops.gemm_nt(st["hn"].t, self.W("f:" + w + ".weight"), st["p"].t[:, j * sg.head_np:], rows,
                        sg.head_n, segs, self.P(w + ".bias"), None, 0.0, 0, self.rng)
        ops.head_reduce_fwd(st["p"].t, B, L, HALO, 2 * sg.head_n, sg.head_n, sg.head_k[0], sg.head_k[1],
                            self.feat[:, sg.feat_off:])

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits: Optional[torch.Tensor], B: int, dropout: bool = True, zero_grads: bool = True,
                 part: int = 0, prefilled: bool = False, bucket_hook=None):
        """Gradient of every parameter into ``self.grads`` given dL/dlogits ([B,out] fp32; None = use the
        engine's own ``dlogits`` buffer written by ``loss_fwd_bwd``).

        part 0 = everything.  The data-parallel step runs it in two parts so that the all-reduce of the first
        gradient bucket overlaps compute (``buckets``): part 1 = output layer, the whole right stream (on the
        second CUDA stream, see ``_fork``) and the left stream down to encoder 1; part 2 = left encoder 0 and the
        Gaussian range encoding."""
        ops, g = self.ops, self.g
        pd = P_DROP if dropout else 0.0
        pf = P_FEAT if dropout else 0.0
        nl = g.left.n_enc
        if part == 2:
            if nl > 1:
                self._backward_stream(0, B, pd, range(0, 1), head=False)
                self._wgrad_join()
            return None
        if not prefilled:                                 # (prefilled: ``forward_body(prefill=True)`` has zeroed both)
            if zero_grads:
                ops.fill_f32(self.grads, 0.0)
            ops.fill_f64(self.red_pool, 0.0)              # parts 1 and 2 use disjoint slices: zeroed once, here
        if dlogits is not None:
            if g.heads == 1:
                self.dlogits[:B, :g.out].copy_(dlogits)
            else:
                self.dlogits[:B].view(B, g.heads, g.cp)[:, :, :g.out].copy_(dlogits.reshape(B, g.heads, g.out))
        ops.alg_scale = 1.0
        ops.dropout_rows(self.dlogits, self.dlogits_a, B, g.ld_out, 0.0, 0, self.rng)      # cast to act dtype
        def w_out():
            for h, (wn, bn) in enumerate(g.output_names()):
                dl = self.dlogits_a[:, h * g.cp:]
                self._alg(2 * B * LY.FEAT * g.out)
                ops.gemm_tn(dl, self.featd, self.G(wn), LY.FEAT, 1, B, g.out, [(0, 0, 0, LY.FEAT)])
                ops.colsum_tokens(dl, B, 1, 0, g.out, self.G(bn))
        self._wgrad(w_out)
        self._alg(2 * B * LY.FEAT * g.out * g.heads)
        ops.gemm_nt(self.dlogits_a, self.W("b:layer_output.weight"), self.dfeatd, B, LY.FEAT,
                    [(0, 0, 0, g.ld_out)], None, None, 0.0, 0, self.rng)
        ops.dropout_rows(self.dfeatd, self.dfeat, B, LY.FEAT, pf, LY.SITE_FEAT, self.rng)
        join = self._fork(lambda: self._backward_stream(1, B, pd, range(0, g.right.n_enc), head=True))
        if bucket_hook is not None and part == 0 and nl > 1:
            # data parallel, ONE launch sequence (one CUDA graph with the all-reduces inside): when the left stream reaches
            # encoder 0 the first gradient bucket is final as soon as the right stream and the weight-gradient streams have
            # drained what they hold -- the all-reduce stream waits for exactly that (events), the data-gradient chain does
            # not wait for anything and goes on into encoder 0
            (lo1, hi1), (lo2, hi2) = self.buckets
            self._backward_stream(0, B, pd, range(1, nl), head=True)
            bucket_hook.start_bucket(self, lo1, hi1, extra_streams=self._busy_streams())
            self._backward_stream(0, B, pd, range(0, 1), head=False)
            join()
            self._wgrad_join()
            bucket_hook.finish(self, lo2, hi2)
            return None
        self._backward_stream(0, B, pd, range(1 if (part == 1 and nl > 1) else 0, nl), head=True)
        join()
        self._wgrad_join()

    @property
    def bucket_split(self) -> int:
        """Arena offset of left encoder 1's first parameter: gradients at or above it (left encoders 1.., the left
        head, the right stream, the output layer: ~3/4 of the bytes) are final after backward part 1, the ones below
        it (Gaussian encoding + left encoder 0) after part 2."""
        if self.g.left.n_enc < 2:
            return 0
        pre = self.g.left.prefix(1)
        split = min(off for k, off in self.arena.offsets.items() if k.startswith(pre))
        # (the heads of the multi-head sibling are registered first: they sit below the split although they are final early)
        early = ("layer_left_gaussian.", self.g.left.prefix(0)) + (("layer_output.",) if self.g.heads > 1 else ())
        assert all((off < split) == k.startswith(early) for k, off in self.arena.offsets.items())
        return split

    @property
    def buckets(self):
        """[(lo, hi) final after part 1, (lo, hi) final after part 2] of the flat gradient arena."""
        return [(self.bucket_split, self.grads.numel()), (0, self.bucket_split)]

    def _backward_stream(self, si: int, B: int, pd: float, encs, head: bool):
        """Backward of one stream (0 = left, 1 = right): [head convs + final norm if ``head``] + the encoders in
        ``encs`` (walked in reverse) [+ the Gaussian encoding when the left stream reaches encoder 0]."""
        ops, g = self.ops, self.g
        sg = g.streams[si]
        st = self.s[sg.name]
        rows, d, Dp, L = sg.rows(B), sg.d, sg.Dp, sg.L
        ops.alg_scale = (L / sg.Lp) * (d / Dp)
        Np = sg.head_np
        # the residual gradient ping-pongs between the two dout buffers, one swap per encoder walked
        flip = (sg.n_enc - 1 - max(encs)) % 2
        dout, dnext = (st["dout"][flip], st["dout"][1 - flip])
        if head:
            ops.head_reduce_bwd(self.dfeat[:, sg.feat_off:], st["p"].t, B, L, HALO, 2 * sg.head_n, sg.head_n,
                                sg.head_k[0], sg.head_k[1], st["dp"].t)
            def w_head():
                for j, k in enumerate(sg.head_k):
                    w = f"layer_{sg.name}_cnn_{j}"
                    self._alg(2 * B * (L - k + 1) * d * sg.head_n * k)
                    ops.gemm_tn(st["dp"].t[:, j * Np:], st["hn"].t, self.G(w + ".weight"), d * k, k, rows, sg.head_n,
                                [(t, 0, t, d) for t in range(k)])
                    ops.colsum_tokens(st["dp"].t[:, j * Np:], B, L, HALO, sg.head_n, self.G(w + ".bias"))
            self._wgrad(w_head)
            dsegs, seg = [], 0
            for j, k in enumerate(sg.head_k):
                dsegs += [(-t, j * Np, (seg + t) * Np, Np) for t in range(k)]
                seg += k
            self._alg(sum(2 * B * (L - k + 1) * d * sg.head_n * k for k in sg.head_k))
            ops.gemm_nt(st["dp"].t, self.W(f"b:layer_{sg.name}_cnn"), st["dhn"].t, rows, d, dsegs, None, None,
                        0.0, 0, self.rng)
            nm = f"layer_{sg.name}_norm."
            ops.layernorm_bwd(st["dhn"].t, st["x_last"].t, self.P(nm + "weight"), st["meanf"], st["rstdf"], None,
                              dout.t, None, 0.0, 0, self.rng, self.G(nm + "weight"), self.G(nm + "bias"),
                              B, L, d, HALO)
        for e in reversed(encs):
          with _Range(f"bwd/{sg.name}/encoder{e}"):
            a, p = st["enc"][e], sg.prefix(e)
            x_in = st["enc"][e - 1]["out"] if e > 0 else st["x0"]
            dz, dtm, dqkv = st["dz"][e], st["dtm"][e], st["dqkv"][e]
            gam, bet = self._bn3(sg, e, "1.weight"), self._bn3(sg, e, "1.bias")
            sb, so = site(si, e, LY.SITE_BRANCH), site(si, e, LY.SITE_SUM)
            ops.bn_act_bwd_reduce(dout.t, a["z"].t, a["bn_mean"], a["bn_invstd"], gam, bet, B, L, d, HALO, 3,
                                  pd, sb, pd, so, self.rng, a["red"], a["dmask"] if pd > 0.0 else None)
            ops.bn_act_bwd_dz(dout.t, a["z"].t, a["bn_mean"], a["bn_invstd"], gam, bet, a["red"], B, L, d,
                              HALO, 3, pd, sb, pd, so, self.rng, dz.t,
                              self._bn3(sg, e, "1.weight", grad=True), self._bn3(sg, e, "1.bias", grad=True),
                              a["dmask"] if pd > 0.0 else None)

            def w_conv(a=a, p=p, dz=dz):
                for j, k in enumerate(sg.kernels):
                    pl = (k - 1) // 2
                    self._alg(2 * B * L * d * d * k)
                    ops.gemm_tn(dz.t[:, j * Dp:], a["s"].t, self.G(f"{p}layer_cnn.{j}.0.weight"), d * k, k,
                                rows, d, [(t - pl, 0, t, d) for t in range(k)])
            self._wgrad(w_conv)
            dsegs, seg = [], 0
            for j, k in enumerate(sg.kernels):
                pl = (k - 1) // 2
                dsegs += [(pl - t, j * Dp, (seg + t) * Dp, Dp) for t in range(k)]
                seg += k
            # the Conv1d biases feed a train-mode BatchNorm: their gradient is identically zero
            self._alg(2 * B * L * d * d * sum(sg.kernels))
            ops.gemm_nt(dz.t, self.W("b:" + p + "layer_cnn"), st["ds"].t, rows, d, dsegs, None, None,
                        0.0, 0, self.rng)
            ops.layernorm_bwd(st["ds"].t, a["t"].t, self.P(p + "layer_norm_1.weight"), a["mean1"], a["rstd1"],
                              dout.t, st["dt"].t, dtm.t, pd, site(si, e, LY.SITE_ATTN), self.rng,
                              self.G(p + "layer_norm_1.weight"), self.G(p + "layer_norm_1.bias"), B, L, d, HALO)
            one = [(0, 0, 0, d)]
            w = p + "layer_attention.out_proj."

            def w_out_proj(a=a, w=w, dtm=dtm):
                self._alg(2 * B * L * d * d)
                ops.gemm_tn(dtm.t, a["o"].t, self.G(w + "weight"), d, 1, rows, d, [(0, 0, 0, sg.dh)], (0, 0), sg.grp)
                ops.colsum_tokens(dtm.t, B, L, HALO, d, self.G(w + "bias"))
            self._wgrad(w_out_proj)
            self._alg(2 * B * L * d * d)
            ops.gemm_nt(dtm.t, self.W("b:" + w + "weight"), st["do"].t, rows, sg.dh, [(0, 0, 0, Dp)], None,
                        None, 0.0, 0, self.rng)
            w = p + "layer_attention.in_proj_"
            # the in_proj bias gradient (column sums of dqkv) is accumulated by the attention backward kernel itself
            ops.attn_bwd(a["qkv"].t, a["o"].t, st["do"].t, dqkv.t, a["lse"], B, L, d, sg.H, sg.hp, HALO,
                         self.G(w + "bias"))

            def w_in_proj(a=a, w=w, dqkv=dqkv):
                self._alg(2 * B * L * d * 3 * d)
                ops.gemm_tn(dqkv.t, a["t0"].t, self.G(w + "weight"), d, 1, rows, sg.ld3, one, sg.grp, (0, 0))
            self._wgrad(w_in_proj)
            self._alg(2 * B * L * d * 3 * d)
            ops.gemm_nt(dqkv.t, self.W("b:" + w + "weight"), st["dt0"].t, rows, d, [(0, 0, 0, sg.ld3)],
                        None, None, 0.0, 0, self.rng)
            ops.layernorm_bwd(st["dt0"].t, x_in.t, self.P(p + "layer_norm_0.weight"), a["mean0"], a["rstd0"],
                              st["dt"].t, dnext.t, None, 0.0, 0, self.rng,
                              self.G(p + "layer_norm_0.weight"), self.G(p + "layer_norm_0.bias"), B, L, d, HALO)
            dout, dnext = dnext, dout
        if si == 0 and 0 in encs:
            gp = "layer_left_gaussian."
            ops.gauss_pe_bwd(dout.t, B, HALO, self.pe_w, self.P(gp + "var_position"), self.P(gp + "var_mu"),
                             self.P(gp + "var_sigma"), self.P(gp + "var_embedding"), L, LY.NUM_GAUSS, g.F,
                             self.dpe_ws, self.G(gp + "var_embedding"), self.G(gp + "var_mu"),
                             self.G(gp + "var_sigma"))

    # ------------------------------------------------------------------ CUDA-graph train body
    def train_body(self, B: int, pos_weight: float, dropout: bool, part: int = 0, bucket_hook=None):
        """repack + forward body + BCE + backward: a fixed launch sequence over static buffers.
        part 1 stops before the left stream's encoder 0 backward, part 2 is that remainder (see ``backward``)."""
        if part != 2:
            # (the operand copies are packed outside: ``ensure_packed`` before the step, ``repack_ahead`` after the optimizer)
            if self.dev.type != "cuda" or not torch.cuda.is_current_stream_capturing():
                self.ensure_packed()
            self.forward_body(B, True, dropout, prefill=self.prefill_on)
            self.loss_fwd_bwd(self.y_static, B, pos_weight)
        self.backward(None, B, dropout=dropout, zero_grads=True, part=part, prefilled=self.prefill_on, bucket_hook=bucket_hook)

    def train_body_graph(self, B: int, pos_weight: float, dropout: bool, part: int = 0, bucket_hook=None):
        """Replays train_body as one CUDA graph (captured on first use for this (B, pos_weight, dropout, part)).  With a
        ``bucket_hook`` the two gradient all-reduces are captured inside the graph (NCCL kernels as graph nodes)."""
        key = (B, float(pos_weight), bool(dropout), part, self.loss_kind, bucket_hook is not None)
        g = self._graphs.get(key)
        if g is None:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            n0 = self.ops.launches
            with torch.cuda.graph(g):
                self.train_body(B, pos_weight, dropout, part, bucket_hook)
            self._graph_launches[key] = self.ops.launches - n0
            self._graphs[key] = g
        else:
            self.ops.launches += self._graph_launches[key]
        g.replay()

    # ------------------------------------------------------------------ loss / optimizer
    def loss_fwd_bwd(self, y: torch.Tensor, B: int, pos_weight: float = 4.0, grad_scale: float = 1.0,
                     want_grad: bool = True) -> torch.Tensor:
        """BCEWithLogitsLoss(pos_weight).mean() of the engine's logits against y [B,out] fp32; writes
        dL/dlogits * grad_scale into the engine's dlogits buffer.  Returns the 1-element loss tensor."""
        if self.loss_kind == "perm_ce":            # multi-head sibling: PermutationMatchingLoss (that_multi_head.py:309-342)
            g = self.g
            self.ops.perm_ce(self.logits, y, B, g.heads, g.out, g.cp, grad_scale, self.loss,
                             self.dlogits if want_grad else None)
        elif self.loss_kind == "smooth_l1":        # sibling head THAT_COUNT_PRED: SmoothL1Loss(beta=1).mean()
            self.ops.smooth_l1(self.logits, y, B, self.g.out, 1.0, grad_scale, self.loss,
                               self.dlogits if want_grad else None)
        else:
            self.ops.bce_logits(self.logits, y, B, self.g.out, pos_weight, grad_scale, self.loss,
                                self.dlogits if want_grad else None)
        return self.loss
