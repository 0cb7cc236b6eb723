"""Static geometry of the THAT train step: token-buffer shapes, the flat parameter arena and the weight
re-layout (pack) table.  Pure Python / torch-CPU metadata -- no kernels here.

Reference shapes: benchmark/wifi_csi/model/that.py:190-245 (layer sizes), :100-139 (Encoder).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

HALO = 2            # zero rows on each side of a sample: covers Conv1d(k<=5, padding="same")
GUARD = 16          # readable rows before/after a token buffer: covers the k=16 head convolution shifts
POOL = 20           # AvgPool1d(kernel_size=20, stride=20), that.py:196,220
NUM_HEAD = 10       # that.py:201,225
NUM_GAUSS = 10      # that.py:37
LEFT_KERNELS = (1, 3, 5)
RIGHT_KERNELS = (1, 2, 3)
NUM_LEFT = 4
NUM_RIGHT = 1
LEFT_HEAD = (128, 8, 16)    # out channels per conv, k0, k1   (that.py:208-214)
RIGHT_HEAD = (16, 2, 4)     # that.py:231-237
FEAT = 2 * LEFT_HEAD[0] + 2 * RIGHT_HEAD[0]   # 288, that.py:245


def ru(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class StreamGeom:
    """One of the two streams: `L` tokens of `d` channels per sample."""
    name: str
    L: int
    d: int
    kernels: Tuple[int, ...]
    n_enc: int
    head_n: int
    head_k: Tuple[int, int]
    feat_off: int
    H: int = NUM_HEAD

    @property
    def Dp(self) -> int:
        return ru(self.d, 16)

    @property
    def Lp(self) -> int:
        return self.L + 2 * HALO

    @property
    def hd(self) -> int:
        return self.d // self.H

    @property
    def hp(self) -> int:
        """Head pitch: heads are stored zero-padded to 16/32/64 channels so every head row is 16-byte aligned."""
        hd = self.hd
        return 16 if hd <= 16 else 32 if hd <= 32 else 64 if hd <= 64 else ru(hd, 16)

    @property
    def dh(self) -> int:
        """Width of a head-padded [*, d] activation (attention output and its gradient)."""
        return self.H * self.hp

    @property
    def ld3(self) -> int:
        """Width of the head-padded q | k | v buffer."""
        return 3 * self.H * self.hp

    @property
    def grp(self):
        """(valid, pad) of the head padding, the csi_grp of include/csi_that.h."""
        return (self.hd, self.hp)

    @property
    def head_np(self) -> int:
        return ru(self.head_n, 16)

    @property
    def nseg_conv(self) -> int:
        return sum(self.kernels)

    @property
    def conv_shifts(self):
        """Row shifts of the fused three-branch forward conv, ascending: the union of the branches' taps (tap t of a size-k
        kernel reads row m + t - (k-1)//2, PyTorch's padding="same")."""
        return sorted({t - (k - 1) // 2 for k in self.kernels for t in range(k)})

    def head_bands(self):
        """(segments, bands) of the fused head convs: tap t (shift t) contributes to the columns of the convs with k > t."""
        Dp, Np, segs, bands = self.Dp, self.head_np, [], []
        ks = list(self.head_k)
        assert ks == sorted(ks), "head convs are stacked by ascending kernel size"
        for t in range(max(ks)):
            first = min(j for j, k in enumerate(ks) if k > t)
            segs.append((t, 0, t * Dp, Dp))
            bands.append((first * Np, len(ks) * Np))
        return segs, bands

    def conv_bands(self):
        """(segments, bands) of csi_gemm_nt_banded for the fused forward conv: segment t = shift conv_shifts[t], contributing to
        the columns of the branches that have that tap."""
        Dp, segs, bands = self.Dp, [], []
        for t, sh in enumerate(self.conv_shifts):
            has = [j for j, k in enumerate(self.kernels) if -((k - 1) // 2) <= sh <= k - 1 - (k - 1) // 2]
            assert has == list(range(has[0], has[-1] + 1)), "branches with a common tap must be adjacent"
            segs.append((sh, 0, t * Dp, Dp))
            bands.append((has[0] * Dp, (has[-1] + 1) * Dp))
        return segs, bands

    def rows(self, B: int) -> int:
        return B * self.Lp

    def prefix(self, e: int) -> str:
        return f"layer_{self.name}_encoder.{e}."


@dataclass
class ModelGeom:
    T: int
    F: int
    out: int
    heads: int = 1          # > 1: the multi-head sibling (model/that_multi_head.py): `heads` Linear(288, out) output layers
    left: StreamGeom = field(init=False)
    right: StreamGeom = field(init=False)

    def __post_init__(self):
        if self.T % POOL:
            raise ValueError("time length must be a multiple of 20 (AvgPool1d(20,20), that.py:196)")
        L = self.T // POOL
        if self.F % NUM_HEAD or L % NUM_HEAD:
            raise ValueError("feature dim and T/20 must be multiples of the 10 attention heads (that.py:201,225)")
        if L < LEFT_HEAD[2] or self.F < RIGHT_HEAD[2]:
            raise ValueError("sequence too short for the head convolutions (that.py:208-237)")
        self.left = StreamGeom("left", L, self.F, LEFT_KERNELS, NUM_LEFT, LEFT_HEAD[0], LEFT_HEAD[1:], 0)
        self.right = StreamGeom("right", self.F, L, RIGHT_KERNELS, NUM_RIGHT, RIGHT_HEAD[0], RIGHT_HEAD[1:],
                                2 * LEFT_HEAD[0])

    @property
    def streams(self):
        return (self.left, self.right)

    @property
    def cp(self) -> int:
        """Column pitch of one output head in the logits buffer."""
        return ru(self.out, 16)

    @property
    def ld_out(self) -> int:
        """Width of the logits buffer: head h occupies columns [h*cp, h*cp + out)."""
        return self.heads * self.cp

    @property
    def n_targets(self) -> int:
        """Label values per sample ([B, out] for THAT, [B, heads, out] for the multi-head sibling)."""
        return self.heads * self.out

    def output_names(self):
        """(weight, bias) parameter names of the output layer(s)."""
        if self.heads == 1:
            return [("layer_output.weight", "layer_output.bias")]
        return [(f"layer_output.{h}.weight", f"layer_output.{h}.bias") for h in range(self.heads)]


def parameter_specs(g: ModelGeom) -> "OrderedDict[str, Tuple[int, ...]]":
    """Trainable + frozen parameter tensors in the reference's registration order (so that
    ``model.parameters()`` and the flat arena enumerate them exactly like the reference module does)."""
    sp: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    L, F = g.left.L, g.F
    if g.heads > 1:                                    # that_multi_head.py:194-196: the ModuleList of heads is registered first
        for w, b in g.output_names():
            sp[w] = (g.out, FEAT)
            sp[b] = (g.out,)
    sp["layer_left_gaussian.var_embedding"] = (NUM_GAUSS, F)
    sp["layer_left_gaussian.var_position"] = (L, NUM_GAUSS)
    sp["layer_left_gaussian.var_mu"] = (1, NUM_GAUSS)
    sp["layer_left_gaussian.var_sigma"] = (1, NUM_GAUSS)

    def enc(prefix, d, kernels):
        sp[prefix + "layer_norm_0.weight"] = (d,)
        sp[prefix + "layer_norm_0.bias"] = (d,)
        sp[prefix + "layer_attention.in_proj_weight"] = (3 * d, d)
        sp[prefix + "layer_attention.in_proj_bias"] = (3 * d,)
        sp[prefix + "layer_attention.out_proj.weight"] = (d, d)
        sp[prefix + "layer_attention.out_proj.bias"] = (d,)
        sp[prefix + "layer_norm_1.weight"] = (d,)
        sp[prefix + "layer_norm_1.bias"] = (d,)
        for j, k in enumerate(kernels):
            sp[f"{prefix}layer_cnn.{j}.0.weight"] = (d, d, k)
            sp[f"{prefix}layer_cnn.{j}.0.bias"] = (d,)
            sp[f"{prefix}layer_cnn.{j}.1.weight"] = (d,)
            sp[f"{prefix}layer_cnn.{j}.1.bias"] = (d,)

    # registration order in that.py:196-245: left encoders, left norm, left cnn_0/1, right ..., output
    for e in range(g.left.n_enc):
        enc(g.left.prefix(e), g.left.d, g.left.kernels)
    sp["layer_left_norm.weight"] = (F,)
    sp["layer_left_norm.bias"] = (F,)
    sp["layer_left_cnn_0.weight"] = (LEFT_HEAD[0], F, LEFT_HEAD[1])
    sp["layer_left_cnn_0.bias"] = (LEFT_HEAD[0],)
    sp["layer_left_cnn_1.weight"] = (LEFT_HEAD[0], F, LEFT_HEAD[2])
    sp["layer_left_cnn_1.bias"] = (LEFT_HEAD[0],)
    for e in range(g.right.n_enc):
        enc(g.right.prefix(e), g.right.d, g.right.kernels)
    sp["layer_right_norm.weight"] = (L,)
    sp["layer_right_norm.bias"] = (L,)
    sp["layer_right_cnn_0.weight"] = (RIGHT_HEAD[0], L, RIGHT_HEAD[1])
    sp["layer_right_cnn_0.bias"] = (RIGHT_HEAD[0],)
    sp["layer_right_cnn_1.weight"] = (RIGHT_HEAD[0], L, RIGHT_HEAD[2])
    sp["layer_right_cnn_1.bias"] = (RIGHT_HEAD[0],)
    if g.heads == 1:
        sp["layer_output.weight"] = (g.out, FEAT)
        sp["layer_output.bias"] = (g.out,)
    return sp


def numel(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


@dataclass
class Arena:
    """Flat fp32 arena: name -> (offset, shape).  Offsets are 4-element aligned (16 B)."""
    offsets: Dict[str, int]
    shapes: Dict[str, Tuple[int, ...]]
    size: int


FROZEN = ("layer_left_gaussian.var_position",)     # nn.Parameter(requires_grad=False), that.py:48


def build_arena(specs) -> Arena:
    """Arena of the TRAINABLE tensors only, so that the flat Adam update never touches frozen ones."""
    off = 0
    offsets, shapes = {}, {}
    for k, shp in specs.items():
        if k in FROZEN:
            continue
        offsets[k] = off
        shapes[k] = tuple(shp)
        off += ru(numel(shp), 4)
    return Arena(offsets, shapes, off)


@dataclass
class PackedMat:
    """A GEMM operand copy of one or more weight tensors in the packed arena."""
    off: int
    rows: int
    ld: int


@dataclass
class PackPlan:
    mats: Dict[str, PackedMat]
    entries: List[tuple]           # csi_pack_entry fields (src, dst, N, C, k, ld, mode, P, seg_base, gn, gc)
    bias_entries: List[tuple]      # same, destination = the fp32 packed-bias arena
    bias_mats: Dict[str, PackedMat]
    bias_size: int
    size: int
    max_elems: int


def build_pack_plan(g: ModelGeom, arena: Arena, fused_conv: bool = True) -> PackPlan:
    """Forward ("f:") and data-gradient ("b:") operand copies of every contraction weight.

    f:<w>   [N, k*Cp]        dst[n, j*Cp + c]              = W[n, c, j]
    b:<grp> [C, nseg*Np]     dst[c, (seg_base+j)*Np + n]   = W[n, c, j]     (segments = taps of all members)
    """
    mats: Dict[str, PackedMat] = {}
    entries = []
    off = 0
    max_elems = 0

    def alloc(name, rows, ld):
        nonlocal off
        mats[name] = PackedMat(off, rows, ld)
        off += ru(rows * ld, 64)
        return mats[name]

    NOG = (0, 0)

    def add(name, N, C, k, dst: PackedMat, mode, P, seg_base, gn=NOG, gc=NOG):
        nonlocal max_elems
        entries.append((arena.offsets[name], dst.off, N, C, k, dst.ld, mode, P, seg_base, gn, gc))
        max_elems = max(max_elems, N * C * k)

    bias_entries, bias_mats, boff = [], {}, 0

    for s in g.streams:
        d, Dp = s.d, s.Dp
        for e in range(s.n_enc):
            p = s.prefix(e)
            # attention projections use the head-padded feature order (q|k|v and the attention output)
            w = p + "layer_attention.in_proj_weight"
            add(w, 3 * d, d, 1, alloc("f:" + w, s.ld3, Dp), 0, Dp, 0, gn=s.grp)
            add(w, 3 * d, d, 1, alloc("b:" + w, d, s.ld3), 1, s.ld3, 0, gn=s.grp)
            bname = p + "layer_attention.in_proj_bias"
            bias_mats[bname] = PackedMat(boff, s.ld3, 1)
            bias_entries.append((arena.offsets[bname], boff, 3 * d, 1, 1, 1, 0, 1, 0, s.grp, NOG))
            boff += ru(s.ld3, 64)
            w = p + "layer_attention.out_proj.weight"
            add(w, d, d, 1, alloc("f:" + w, d, s.dh), 0, s.dh, 0, gc=s.grp)
            add(w, d, d, 1, alloc("b:" + w, s.dh, Dp), 1, Dp, 0, gc=s.grp)
            bmat = alloc("b:" + p + "layer_cnn", d, s.nseg_conv * Dp)
            # fused forward operand of the three branches (csi_gemm_nt_banded): rows j*Dp.. = branch j, column block t =
            # row shift conv_shifts[t]; the taps a branch does not have stay zero (the buffer is zero-initialised)
            shifts = s.conv_shifts
            fused = alloc("f:" + p + "layer_cnn", len(s.kernels) * Dp, len(shifts) * Dp)
            seg = 0
            for j, k in enumerate(s.kernels):
                w = f"{p}layer_cnn.{j}.0.weight"
                if not fused_conv:                                         # per-branch operands: only the three-launch path reads them
                    add(w, d, d, k, alloc("f:" + w, d, k * Dp), 0, Dp, 0)
                first = shifts.index(-((k - 1) // 2))                      # column block of the branch's first tap
                add(w, d, d, k, PackedMat(fused.off + j * Dp * fused.ld + first * Dp, d, fused.ld), 0, Dp, 0)
                add(w, d, d, k, bmat, 1, Dp, seg)
                seg += k
        Np = s.head_np
        bmat = alloc(f"b:layer_{s.name}_cnn", d, sum(s.head_k) * Np)
        seg = 0
        # fused forward operand of the two head convs (csi_gemm_nt_banded): rows j*Np.. = head conv j, column block t = tap t
        # (valid convolutions: both start at shift 0); + their biases side by side at the same pitch
        hfused = alloc(f"f:layer_{s.name}_cnn", len(s.head_k) * Np, max(s.head_k) * Dp)
        bias_mats[f"layer_{s.name}_cnn.bias"] = PackedMat(boff, len(s.head_k) * Np, 1)
        for j, k in enumerate(s.head_k):
            w = f"layer_{s.name}_cnn_{j}.weight"
            if not fused_conv:
                add(w, s.head_n, d, k, alloc("f:" + w, s.head_n, k * Dp), 0, Dp, 0)
            add(w, s.head_n, d, k, PackedMat(hfused.off + j * Np * hfused.ld, s.head_n, hfused.ld), 0, Dp, 0)
            bias_entries.append((arena.offsets[f"layer_{s.name}_cnn_{j}.bias"], boff + j * Np, s.head_n, 1, 1, 1, 0, 1, 0, NOG, NOG))
            add(w, s.head_n, d, k, bmat, 1, Np, seg)
            seg += k
        boff += ru(len(s.head_k) * Np, 64)
    # output layer(s): one forward operand [heads*cp, 288] (head h = rows h*cp..), one data-gradient operand [288, heads*cp]
    # (head h = segment h of width cp) and -- for several heads -- one packed bias vector with the same pitch
    fmat = alloc("f:layer_output.weight", g.ld_out if g.heads > 1 else g.out, FEAT)
    bmat = alloc("b:layer_output.weight", FEAT, g.ld_out)
    for h, (w, bname) in enumerate(g.output_names()):
        add(w, g.out, FEAT, 1, PackedMat(fmat.off + h * g.cp * FEAT, g.out, FEAT), 0, FEAT, 0)
        add(w, g.out, FEAT, 1, bmat, 1, g.cp if g.heads > 1 else g.ld_out, h)
        if g.heads > 1:
            if h == 0:
                bias_mats["layer_output.bias"] = PackedMat(boff, g.ld_out, 1)
            bias_entries.append((arena.offsets[bname], boff + h * g.cp, g.out, 1, 1, 1, 0, 1, 0, NOG, NOG))
    if g.heads > 1:
        boff += ru(g.ld_out, 64)
    return PackPlan(mats=mats, entries=entries, size=off, max_elems=max_elems, bias_entries=bias_entries,
                    bias_mats=bias_mats, bias_size=boff)


# dropout site ids (unique per mask): stream*1000 + encoder*16 + kind
SITE_ATTN, SITE_BRANCH, SITE_SUM, SITE_FEAT, SITE_AUG, SITE_AUG_SCALE = 0, 1, 4, 9000, 9001, 9002


def site(stream_idx: int, enc: int, kind: int) -> int:
    return (stream_idx + 1) * 1000 + enc * 16 + kind
