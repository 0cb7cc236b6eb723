"""GPU parity of every C-ABI kernel against the torch mirror (tests/mirror_ops.py) on the same inputs.
Runs through multi_modal_csi_b200.ops.NativeOps, i.e. through the ctypes boundary of libcsi_that.so."""
import math

import numpy as np
import pytest
import torch

from mirror_ops import MirrorOps

pytestmark = pytest.mark.gpu

GUARD = 16
HALO = 2


def ru(x, m):
    return (x + m - 1) // m * m


@pytest.fixture(scope="module")
def ops():
    from multi_modal_csi_b200.ops import NativeOps
    return NativeOps(torch.device("cuda", 0))


@pytest.fixture(scope="module")
def mir():
    return MirrorOps("cuda")


def tokbuf(B, L, ld, dtype, fill=None, gen=None, ncols=None):
    """Token buffer with zero halo/guard rows; valid rows filled with N(0,1)*fill in the first ncols columns."""
    Lp = L + 2 * HALO
    rows = B * Lp
    full = torch.zeros(rows + 2 * GUARD, ld, dtype=dtype, device="cuda")
    body = full[GUARD:GUARD + rows]
    if fill is not None:
        ncols = ld if ncols is None else ncols
        v = torch.randn(B, L, ncols, device="cuda", generator=gen) * fill
        body.view(B, Lp, ld)[:, HALO:HALO + L, :ncols] = v.to(dtype)
    return full, body


def valid(body, B, L, ncols):
    Lp = L + 2 * HALO
    return body.view(B, Lp, -1)[:, HALO:HALO + L, :ncols].float()


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 1e-2}
SHAPES = [(3, 20, 30), (2, 150, 270), (2, 270, 150), (1, 540, 150)]


def gen(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


# ------------------------------------------------------------------------------------------------ input stage
@pytest.mark.parametrize("B,T,F", [(3, 400, 30), (2, 3000, 270), (1, 3000, 540)])
def test_pool_dual_dense(ops, mir, B, T, F):
    g = gen(1)
    L = T // 20
    x = torch.rand(B, T, F, device="cuda", generator=g) * 20
    pe = torch.randn(L, ru(F, 16), device="cuda", generator=g)
    outs = []
    for o in (ops, mir):
        _, left = tokbuf(B, L, ru(F, 16), torch.float32)
        _, right = tokbuf(B, F, ru(L, 16), torch.float32)
        o.pool_dual(x, None, None, B, T, F, pe, left, right, HALO, 0, None)
        outs.append((left, right))
    assert relerr(outs[0][0], outs[1][0]) < 1e-6
    assert relerr(outs[0][1], outs[1][1]) < 1e-6
    ref = torch.nn.functional.avg_pool1d(x.transpose(1, 2), 20, 20)          # [B,F,L]
    assert relerr(valid(outs[0][1], B, F, L), ref) < 1e-6


@pytest.mark.parametrize("T,F,lens", [(400, 30, [400, 371, 1, 260]), (3000, 270, [3000, 2871, 1, 1777]),
                                      (3000, 540, [2999, 3000, 20, 1501])])
def test_pool_dual_ragged_front_pad(ops, mir, T, F, lens):
    """Odd pad lengths put the first valid row of a token at an 8-byte (not 16-byte) aligned address when F/2 is odd: the
    staged kernel moves the unaligned head/tail of each run with ordinary loads."""
    B = 4
    L = T // 20
    g = gen(2)
    lens = torch.tensor(lens, dtype=torch.int32)
    offs = torch.zeros(B, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens[:-1].long() * F, 0)
    arena = torch.rand(int((lens.long() * F).sum()), device="cuda", generator=g) * 20
    outs = []
    for o in (ops, mir):
        _, left = tokbuf(B, L, ru(F, 16), torch.float32)
        _, right = tokbuf(B, F, ru(L, 16), torch.float32)
        o.pool_dual(arena, offs.cuda(), lens.cuda(), B, T, F, None, left, right, HALO, 0, None)
        outs.append((left, right))
    assert relerr(outs[0][0], outs[1][0]) < 1e-6
    assert relerr(outs[0][1], outs[1][1]) < 1e-6
    # the front of short samples is zero (load_data.py:70-72 pads in FRONT)
    lv = valid(outs[0][0], B, L, F)
    short = int(lens.argmin())
    first = L - (int(lens[short]) + 19) // 20
    assert float(lv[short, :first].abs().max()) == 0.0 and float(lv[short, first].abs().max()) > 0.0


def test_pool_dual_augment_statistics(ops):
    B, T, F = 8, 3000, 270
    L = T // 20
    x = torch.full((B, T, F), 10.0, device="cuda")
    rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
    _, left = tokbuf(B, L, ru(F, 16), torch.float32)
    _, right = tokbuf(B, F, ru(L, 16), torch.float32)
    ops.pool_dual(x, None, None, B, T, F, None, left, right, HALO, 1, rng)
    lv = valid(left, B, L, F)                                           # mean over 20 of (10+.1n)*s*m
    per_sample = lv.mean(dim=(1, 2)) / 10.0 / 0.96                      # ~ scale_b in [0.9, 1.1)
    assert float(per_sample.min()) > 0.89 and float(per_sample.max()) < 1.11
    assert float(per_sample.std()) > 0.01                               # scales differ per sample
    # pooled variance: keep-mask Bernoulli(.96) on a constant 10*s dominates; noise adds 0.01*s^2/20
    s = per_sample.view(B, 1, 1)
    resid = lv / s
    expect_var = (100.0 * 0.96 * 0.04 + 0.01 * 0.96) / 20.0
    assert abs(float(resid.var()) / expect_var - 1.0) < 0.05
    assert relerr(valid(right, B, F, L), lv.transpose(1, 2)) < 1e-6     # both streams see the same pooled tensor
    # a different step gives a different draw, the same step the same draw
    _, left2 = tokbuf(B, L, ru(F, 16), torch.float32)
    ops.pool_dual(x, None, None, B, T, F, None, left2, right, HALO, 1, rng)
    assert torch.equal(left, left2)
    rng2 = torch.tensor([1234, 8], dtype=torch.int64, device="cuda")
    ops.pool_dual(x, None, None, B, T, F, None, left2, right, HALO, 1, rng2)
    assert not torch.equal(left, left2)


def test_gauss_pe(ops, mir):
    L, K, F = 150, 10, 270
    g = gen(3)
    pos = torch.arange(0.0, L, device="cuda").unsqueeze(1).repeat(1, K).contiguous()
    mu = torch.arange(0.0, L, L / K, device="cuda") + torch.randn(K, device="cuda", generator=g)
    sigma = 50 + 5 * torch.randn(K, device="cuda", generator=g)
    emb = torch.randn(K, F, device="cuda", generator=g) * 0.1
    res = []
    for o in (ops, mir):
        w = torch.zeros(L, K, device="cuda")
        pe = torch.zeros(L, ru(F, 16), device="cuda")
        o.gauss_pe_fwd(pos, mu, sigma, emb, L, K, F, w, pe)
        B = 5
        _, dleft = tokbuf(B, L, ru(F, 16), torch.float32, fill=1.0, gen=gen(4), ncols=F)
        ws = torch.zeros(L, ru(F, 16), device="cuda")
        demb, dmu, dsg = torch.zeros(K, F, device="cuda"), torch.zeros(K, device="cuda"), torch.zeros(K, device="cuda")
        o.gauss_pe_bwd(dleft, B, HALO, w, pos, mu, sigma, emb, L, K, F, ws, demb, dmu, dsg)
        res.append((w, pe, demb, dmu, dsg))
    for a, b in zip(*res):
        assert relerr(a, b) < 1e-4


# ------------------------------------------------------------------------------------------------ layernorm
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,L,d", SHAPES + [(2, 150, 540)])
def test_layernorm(ops, mir, dt, B, L, d):
    ld = ru(d, 16)
    g = gen(5)
    _, x = tokbuf(B, L, ld, torch.float32, fill=3.0, gen=g, ncols=d)
    x += 1.5 * (x != 0)
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    _, dy = tokbuf(B, L, ld, dt, fill=1.0, gen=g, ncols=d)
    _, dres = tokbuf(B, L, ld, torch.float32, fill=1.0, gen=g, ncols=d)
    res = []
    for o in (ops, mir):
        _, y = tokbuf(B, L, ld, dt)
        rows = B * (L + 2 * HALO)
        mean, rstd = torch.zeros(rows, device="cuda"), torch.zeros(rows, device="cuda")
        o.layernorm_fwd(x, gamma, beta, y, mean, rstd, B, L, d, HALO, 1e-6)
        _, dx = tokbuf(B, L, ld, torch.float32)
        _, dxm = tokbuf(B, L, ld, dt)
        dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        o.layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dxm, 0.0, 0, None, dg, db, B, L, d, HALO)
        res.append((y.float(), mean, rstd, dx, dxm.float(), dg, db))
    tol = TOL[dt]
    for i, (a, b) in enumerate(zip(*res)):
        assert relerr(a, b) < tol, i
    ref = torch.nn.functional.layer_norm(valid(x, B, L, d), (d,), gamma, beta, 1e-6)
    assert relerr(valid(res[0][0], B, L, d), ref) < tol
    assert float(res[0][0][:, d:].abs().max()) == 0.0          # pad columns stay zero


# ------------------------------------------------------------------------------------------------ GEMMs
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("simt", [True, False])
def test_gemm_nt_linear_and_conv(ops, mir, dt, simt):
    ops.set_force_simt(simt)
    try:
        B, L, d, N = 3, 150, 270, 270
        Dp = ru(d, 16)
        g = gen(6)
        _, A = tokbuf(B, L, Dp, dt, fill=1.0, gen=g, ncols=d)
        rows = B * (L + 2 * HALO)
        bias = torch.randn(N, device="cuda", generator=g)
        _, res = tokbuf(B, L, Dp, torch.float32, fill=1.0, gen=g, ncols=d)
        for k, cdt in ((1, torch.float32), (3, dt), (5, dt)):
            W = torch.zeros(N, k * Dp, dtype=dt, device="cuda")
            for j in range(k):
                W[:, j * Dp:j * Dp + d] = (torch.randn(N, d, device="cuda", generator=g) / math.sqrt(d * k)).to(dt)
            pl = (k - 1) // 2
            segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
            outs = []
            for o in (ops, mir):
                _, Cm = tokbuf(B, L, ru(N, 16), cdt)
                o.gemm_nt(A, W, Cm, rows, N, segs, bias, res if k == 1 else None, 0.0, 0, None)
                outs.append(Cm.float())
            assert relerr(outs[0], outs[1]) < (3e-6 if dt == torch.float32 and cdt == torch.float32 else 6e-3), (k, cdt)
            if k == 3:                                                     # against torch's own Conv1d
                torch.backends.cudnn.allow_tf32 = False
                xin = valid(A, B, L, d).transpose(1, 2)
                w3 =torch.stack([W[:, j * Dp:j * Dp + d].float() for j in range(k)], dim=-1)
                ref = torch.nn.functional.conv1d(xin, w3, bias, padding=1).transpose(1, 2)
                assert relerr(valid(outs[0], B, L, N), ref) < (1e-5 if dt == torch.float32 else 1e-2)
    finally:
        ops.set_force_simt(False)


@pytest.mark.parametrize("N,k,f32out", [(270, 5, False), (270, 1, True), (128, 8, False), (150, 3, False)])
def test_gemm_nt_full_batch_tail_wave(ops, mir, N, k, f32out):
    """B=256 token buffers: 308 row tiles on 148 SMs.  The tiles of the last, partly filled wave are cut into 64-column
    pieces (gemm_tc3.cu, Nt3Params.nfull/npiece); rows of those tiles must come out exactly like the others."""
    dt = torch.bfloat16
    B, L, d = 256, 150, 270
    Dp = ru(d, 16)
    g = gen(16)
    _, A = tokbuf(B, L, Dp, dt, fill=1.0, gen=g, ncols=d)
    rows = B * (L + 2 * HALO)
    bias = torch.randn(N, device="cuda", generator=g) if f32out else None
    res = tokbuf(B, L, ru(N, 16), torch.float32, fill=1.0, gen=g, ncols=N)[1] if f32out else None
    W = torch.zeros(N, k * Dp, dtype=dt, device="cuda")
    for j in range(k):
        W[:, j * Dp:j * Dp + d] = (torch.randn(N, d, device="cuda", generator=g) / math.sqrt(d * k)).to(dt)
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
    outs = []
    for o in (ops, mir):
        _, Cm = tokbuf(B, L, ru(N, 16), torch.float32 if f32out else dt)
        o.gemm_nt(A, W, Cm, rows, N, segs, bias, res, 0.0, 0, None)
        outs.append(Cm.float())
    assert relerr(outs[0], outs[1]) < 6e-3
    tail = outs[0][296 * 128:], outs[1][296 * 128:]                    # rows of the split tiles
    assert relerr(*tail) < 6e-3
    if outs[0].shape[1] > N:
        assert float(outs[0][:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("M,N,K", [(256, 54, 288), (37, 12, 288), (1000, 810, 272), (300, 16, 160)])
def test_gemm_nt_plain(ops, mir, dt, M, N, K):
    g = gen(7)
    A = torch.randn(M, K, device="cuda", generator=g).to(dt)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dt)
    bias = torch.randn(N, device="cuda", generator=g)
    outs = []
    for o in (ops, mir):
        Cm = torch.zeros(M, ru(N, 16), device="cuda")
        o.gemm_nt(A, W, Cm, M, N, [(0, 0, 0, K)], bias, None, 0.0, 0, None)
        outs.append(Cm)
    assert relerr(outs[0], outs[1]) < (3e-6 if dt == torch.float32 else 2e-5)
    if outs[0].shape[1] > N:
        assert float(outs[0][:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("dt", DT)
def test_gemm_tn_weight_gradients(ops, mir, dt):
    B, L, d, N, k = 3, 150, 270, 128, 8
    Dp = ru(d, 16)
    g = gen(8)
    rows = B * (L + 2 * HALO)
    _, dY = tokbuf(B, L, ru(N, 16), dt, fill=1.0, gen=g, ncols=N)
    _, X = tokbuf(B, L, Dp, dt, fill=1.0, gen=g, ncols=d)
    outs = []
    for o in (ops, mir):
        gw = torch.zeros(N * d * k, device="cuda")
        o.gemm_tn(dY, X, gw, d * k, k, rows, N, [(t, 0, t, d) for t in range(k)])
        gl = torch.zeros(N * d, device="cuda")
        o.gemm_tn(dY, X, gl, d, 1, rows, N, [(0, 0, 0, d)])
        outs.append((gw, gl))
    tol = 2e-5 if dt == torch.float32 else 2e-5
    assert relerr(outs[0][0], outs[1][0]) < tol
    assert relerr(outs[0][1], outs[1][1]) < tol
    # independent check of the conv layout: autograd of a valid Conv1d whose output gradient is dY
    xin = valid(X, B, L, d).transpose(1, 2).clone().requires_grad_(False)
    w = torch.zeros(N, d, k, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv1d(xin, w)                                    # [B,N,L-k+1]
    gy = valid(dY, B, L, N).transpose(1, 2)[:, :, :L - k + 1]
    y.backward(gy)
    # rows t > L-k of dY also contribute in the kernel (they read the next rows); zero them for this check
    _, dY2 = tokbuf(B, L, ru(N, 16), dt)
    dY2.view(B, L + 2 * HALO, -1)[:, HALO:HALO + L - k + 1] = dY.view(B, L + 2 * HALO, -1)[:, HALO:HALO + L - k + 1]
    gw2 = torch.zeros(N * d * k, device="cuda")
    ops.gemm_tn(dY2, X, gw2, d * k, k, rows, N, [(t, 0, t, d) for t in range(k)])
    assert relerr(gw2.view(N, d, k), w.grad) < (1e-5 if dt == torch.float32 else 1e-5)
    # head-padded operands: rows / columns of padding are skipped and the result lands in the compact layout
    H, hd, hp = 10, 27, 32
    A2 = padded_heads(B, L, H, hd, hp, 1, dt, g)                     # [rows, 320]: 10 heads x (27 valid + 5 pad)
    outs = []
    for o in (ops, mir):
        c1 = torch.zeros(H * hd * d, device="cuda")
        o.gemm_tn(A2, X, c1, d, 1, rows, H * hp, [(0, 0, 0, d)], (hd, hp), (0, 0))          # rows compacted
        c2 = torch.zeros(d * H * hd, device="cuda")
        o.gemm_tn(X, A2, c2, H * hd, 1, rows, d, [(0, 0, 0, H * hp)], (0, 0), (hd, hp))     # columns compacted
        outs.append((c1, c2))
    assert relerr(outs[0][0], outs[1][0]) < 2e-5 and relerr(outs[0][1], outs[1][1]) < 2e-5
    dense = valid(A2, B, L, H * hp).reshape(-1, H, hp)[..., :hd].reshape(-1, H * hd)
    assert relerr(outs[0][0].view(H * hd, d), dense.t() @ valid(X, B, L, d).reshape(-1, d)) < 2e-5


@pytest.mark.parametrize("dt", DT)
def test_colsum(ops, mir, dt):
    B, L, n = 3, 150, 810
    g = gen(9)
    _, A = tokbuf(B, L, ru(n, 16), dt, fill=1.0, gen=g, ncols=n)
    A[0:2] = 7.0                                                # halo rows must be ignored
    outs = []
    for o in (ops, mir):
        out = torch.zeros(n, device="cuda")
        o.colsum_tokens(A, B, L, HALO, n, out)
        outs.append(out)
    assert relerr(outs[0], outs[1]) < 1e-5
    x = torch.randn(37, 16, device="cuda", generator=g)
    out = torch.zeros(13, device="cuda")
    ops.colsum_tokens(x, 37, 1, 0, 13, out)
    assert relerr(out, x[:, :13].sum(0)) < 1e-5
    # head-padded columns are compacted: 4 heads of 3 valid + 1 padding column
    out = torch.zeros(12, device="cuda")
    ops.colsum_tokens(x, 37, 1, 0, 16, out, (3, 4))
    assert relerr(out, x.view(37, 4, 4)[:, :, :3].reshape(37, 12).sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ attention
def head_pitch(hd):
    return 16 if hd <= 16 else 32 if hd <= 32 else 64


def padded_heads(B, L, H, hd, hp, nw, dt, g):
    """Token buffer [rows, nw*H*hp] with N(0,1) in the hd valid columns of every head and zeros in the padding."""
    _, buf = tokbuf(B, L, nw * H * hp, dt)
    v = torch.randn(B, L, nw * H, hp, device="cuda", generator=g)
    v[..., hd:] = 0
    buf.view(B, L + 2 * HALO, -1)[:, HALO:HALO + L] = v.reshape(B, L, nw * H * hp).to(dt)
    return buf


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,L,d", [(2, 20, 30), (2, 30, 20), (2, 150, 270), (2, 270, 150), (1, 150, 540), (1, 540, 150)])
def test_attention(ops, mir, dt, B, L, d):
    H = 10
    hd = d // H
    hp = head_pitch(hd)
    g = gen(10)
    qkv = padded_heads(B, L, H, hd, hp, 3, dt, g)
    do = padded_heads(B, L, H, hd, hp, 1, dt, g)
    res = []
    for o in (ops, mir):
        _, out = tokbuf(B, L, H * hp, dt)
        lse = torch.zeros(B * H * L, device="cuda")
        o.attn_fwd(qkv, out, lse, B, L, d, H, hp, HALO)
        _, dqkv = tokbuf(B, L, 3 * H * hp, dt)
        dbias = torch.full((3 * d,), 0.5, device="cuda")            # accumulated into, not overwritten
        o.attn_bwd(qkv, out, do, dqkv, lse, B, L, d, H, hp, HALO, dbias)
        res.append((out.float(), lse, dqkv.float(), dbias))
    tol = 2e-5 if dt == torch.float32 else 1.5e-2
    for i, (a, b) in enumerate(zip(*res)):
        assert relerr(a, b) < tol, i
    # the fused in_proj bias gradient is the column sum of the stored dqkv (compact channel order)
    cs = valid(res[0][2], B, L, 3 * H * hp).reshape(B * L, 3, H, hp)[..., :hd].sum(0).reshape(-1) + 0.5
    assert relerr(res[0][3], cs) < 1e-5
    # padding columns of the outputs are zero
    assert float(valid(res[0][0], B, L, H * hp).reshape(B, L, H, hp)[..., hd:].abs().max()) == 0.0
    assert float(valid(res[0][2], B, L, 3 * H * hp).reshape(B, L, 3 * H, hp)[..., hd:].abs().max()) == 0.0
    # against torch's scaled_dot_product_attention
    t = valid(qkv, B, L, 3 * H * hp).reshape(B, L, 3, H, hp)[..., :hd]
    q, k, v = [t[:, :, w].transpose(1, 2) for w in range(3)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2)          # [B,L,H,hd]
    got = valid(res[0][0], B, L, H * hp).reshape(B, L, H, hp)[..., :hd]
    assert relerr(got, ref) < (1e-5 if dt == torch.float32 else 1e-2)


# ------------------------------------------------------------------------------------------------ batchnorm block
@pytest.mark.parametrize("dt", DT)
def test_bn_block_stored_dropout_masks(ops, dt):
    """The keep bits bn_act_fwd stores are exactly the Philox decisions: the backward kernels give bit-identical results
    whether they read the stored words or regenerate them, and the drop rate / keep pattern is the one of the forward."""
    B, L, d = 3, 150, 270
    Dp = ru(d, 16)
    g = gen(12)
    _, z = tokbuf(B, L, 3 * Dp, dt, fill=1.0, gen=g, ncols=3 * Dp)
    _, t = tokbuf(B, L, Dp, torch.float32, fill=1.0, gen=g, ncols=d)
    _, dout = tokbuf(B, L, Dp, torch.float32, fill=1.0, gen=g, ncols=d)
    gam = [1 + 0.1 * torch.randn(d, device="cuda", generator=g) for _ in range(3)]
    bet = [0.1 * torch.randn(d, device="cuda", generator=g) for _ in range(3)]
    mean, inv = 0.1 * torch.randn(3 * Dp, device="cuda", generator=g), 1 + 0.1 * torch.rand(3 * Dp, device="cuda", generator=g)
    rng = torch.tensor([77, 5], dtype=torch.int64, device="cuda")
    rows = z.shape[0]
    outs = []
    for use_masks in (False, True):
        masks = torch.zeros(rows * (Dp // 8), dtype=torch.int32, device="cuda") if use_masks else None
        _, out = tokbuf(B, L, Dp, torch.float32)
        ops.bn_act_fwd(z, mean, inv, gam, bet, t, out, B, L, d, HALO, 3, 0.1, 40, 0.1, 50, rng, masks)
        red = torch.zeros(2 * 3 * Dp, dtype=torch.float64, device="cuda")
        ops.bn_act_bwd_reduce(dout, z, mean, inv, gam, bet, B, L, d, HALO, 3, 0.1, 40, 0.1, 50, rng, red, masks)
        _, dz = tokbuf(B, L, 3 * Dp, dt)
        dg = [torch.zeros(d, device="cuda") for _ in range(3)]
        db = [torch.zeros(d, device="cuda") for _ in range(3)]
        ops.bn_act_bwd_dz(dout, z, mean, inv, gam, bet, red, B, L, d, HALO, 3, 0.1, 40, 0.1, 50, rng, dz, dg, db, masks)
        outs.append((out.clone(), red.clone(), dz.float().clone(), torch.cat(dg), torch.cat(db), masks))
    for a, b in zip(outs[0][:5], outs[1][:5]):
        assert torch.equal(a, b)
    m = outs[1][5].view(B, L + 2 * HALO, Dp // 8)[:, HALO:HALO + L, :d // 8]
    bits = torch.stack([(m >> s) & 1 for s in range(32)], -1).float()
    assert abs(bits.mean().item() - 0.9) < 0.01


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,L,d", SHAPES)
def test_bn_block(ops, mir, dt, B, L, d):
    Dp = ru(d, 16)
    g = gen(11)
    _, z = tokbuf(B, L, 3 * Dp, dt)
    zb = z.view(B, L + 2 * HALO, 3 * Dp)
    for br in range(3):
        zb[:, HALO:HALO + L, br * Dp:br * Dp + d] = (torch.randn(B, L, d, device="cuda", generator=g) * (br + 1) + 0.3 * br).to(dt)
    _, t = tokbuf(B, L, Dp, torch.float32, fill=1.0, gen=g, ncols=d)
    _, dout = tokbuf(B, L, Dp, torch.float32, fill=1.0, gen=g, ncols=d)
    gam = [1 + 0.1 * torch.randn(d, device="cuda", generator=g) for _ in range(3)]
    bet = [0.1 * torch.randn(d, device="cuda", generator=g) for _ in range(3)]
    cb = [0.1 * torch.randn(d, device="cuda", generator=g) for _ in range(3)]
    res = []
    for o in (ops, mir):
        rm = [torch.zeros(d, device="cuda") for _ in range(3)]
        rv = [torch.ones(d, device="cuda") for _ in range(3)]
        nbt = [torch.zeros((), dtype=torch.long, device="cuda") for _ in range(3)]
        sums = torch.zeros(2 * 3 * Dp, dtype=torch.float64, device="cuda")
        mean, inv = torch.zeros(3 * Dp, device="cuda"), torch.zeros(3 * Dp, device="cuda")
        o.bn_stats(z, B, L, HALO, 3 * Dp, sums)
        o.bn_finalize(sums, Dp, d, 3, B * L, cb, rm, rv, nbt, 0.1, 1e-5, mean, inv)
        _, out = tokbuf(B, L, Dp, torch.float32)
        o.bn_act_fwd(z, mean, inv, gam, bet, t, out, B, L, d, HALO, 3, 0.0, 0, 0.0, 0, None)
        red = torch.zeros(2 * 3 * Dp, dtype=torch.float64, device="cuda")
        o.bn_act_bwd_reduce(dout, z, mean, inv, gam, bet, B, L, d, HALO, 3, 0.0, 0, 0.0, 0, None, red)
        _, dz = tokbuf(B, L, 3 * Dp, dt)
        dg = [torch.zeros(d, device="cuda") for _ in range(3)]
        db = [torch.zeros(d, device="cuda") for _ in range(3)]
        o.bn_act_bwd_dz(dout, z, mean, inv, gam, bet, red, B, L, d, HALO, 3, 0.0, 0, 0.0, 0, None, dz, dg, db)
        me, ie = torch.zeros(3 * Dp, device="cuda"), torch.zeros(3 * Dp, device="cuda")
        o.bn_eval_prepare(Dp, d, 3, cb, rm, rv, 1e-5, me, ie)
        res.append((mean, inv, out, dz.float(), torch.cat(dg), torch.cat(db), torch.cat(rm), torch.cat(rv),
                    torch.stack(nbt).float(), me, ie))
    tol = 3e-5 if dt == torch.float32 else 1e-2
    for i, (a, b) in enumerate(zip(*res)):
        assert relerr(a, b) < tol, i
    # against torch BatchNorm1d autograd
    zz = [valid(z, B, L, 3 * Dp)[..., br * Dp:br * Dp + d].transpose(1, 2).clone().requires_grad_(True) for br in range(3)]
    acc = 0
    for br in range(3):
        y = torch.nn.functional.batch_norm(zz[br], None, None, gam[br], bet[br], True, 0.1, 1e-5)
        acc = acc + torch.nn.functional.leaky_relu(y, 0.01)
    outr = (acc / 3).transpose(1, 2) + valid(t, B, L, d)
    assert relerr(valid(res[0][2], B, L, d), outr) < tol
    outr.backward(valid(dout, B, L, d))
    for br in range(3):
        got = valid(res[0][3], B, L, 3 * Dp)[..., br * Dp:br * Dp + d]
        assert relerr(got, zz[br].grad.transpose(1, 2)) < (1e-4 if dt == torch.float32 else 2e-2)


# ------------------------------------------------------------------------------------------------ heads / loss / adam / pack
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,L,N,k0,k1", [(3, 150, 128, 8, 16), (3, 270, 16, 2, 4), (2, 20, 128, 8, 16)])
def test_head_reduce(ops, mir, dt, B, L, N, k0, k1):
    g = gen(12)
    _, p = tokbuf(B, L, 2 * N, dt, fill=1.0, gen=g)
    dfeat = torch.randn(B, 288, device="cuda", generator=g)
    res = []
    for o in (ops, mir):
        feat = torch.zeros(B, 288, device="cuda")
        o.head_reduce_fwd(p, B, L, HALO, 2 * N, N, k0, k1, feat[:, 32:])
        _, dp = tokbuf(B, L, 2 * N, dt)
        o.head_reduce_bwd(dfeat[:, 32:], p, B, L, HALO, 2 * N, N, k0, k1, dp)
        res.append((feat, dp.float()))
    for a, b in zip(*res):
        assert relerr(a, b) < 1e-5
    assert float(res[0][0][:, :32].abs().max()) == 0.0


def test_bce_and_dropout_and_adam(ops, mir):
    g = gen(13)
    B, out = 37, 54
    z = torch.zeros(B, 64, device="cuda")
    z[:, :out] = torch.randn(B, out, device="cuda", generator=g) * 8
    y = (torch.rand(B, out, device="cuda", generator=g) < 0.15).float()
    loss, dz = torch.zeros(1, device="cuda"), torch.zeros(B, 64, device="cuda")
    ops.bce_logits(z, y, B, out, 4.0, 1.0, loss, dz)
    zr = z[:, :out].clone().requires_grad_(True)
    lr_ = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))(zr, y)
    lr_.backward()
    assert abs(loss.item() - lr_.item()) < 1e-5 * max(1, abs(lr_.item()))
    assert relerr(dz[:, :out], zr.grad) < 1e-5
    # dropout: keep rate, scaling, determinism per (seed, step, site)
    rng = torch.tensor([99, 3], dtype=torch.int64, device="cuda")
    x = torch.ones(512, 288, device="cuda")
    o1, o2 = torch.zeros(512, 288, device="cuda"), torch.zeros(512, 288, dtype=torch.bfloat16, device="cuda")
    ops.dropout_rows(x, o1, 512, 288, 0.5, 9000, rng)
    ops.dropout_rows(x, o2, 512, 288, 0.5, 9000, rng)
    assert torch.equal(o1, o2.float())
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - 0.5) < 0.01 and set(o1.unique().tolist()) == {0.0, 2.0}
    ops.dropout_rows(x, o2, 512, 288, 0.5, 9001, rng)
    assert not torch.equal(o1, o2.float())
    o3 = torch.zeros(512, 288, device="cuda")
    ops.dropout_rows(x, o3, 512, 288, 0.1, 5, rng)
    assert abs((o3 != 0).float().mean().item() - 0.9) < 0.01
    assert abs(o3.max().item() - 1 / 0.9) < 1e-6
    # adam vs torch.optim.Adam (coupled L2), 3 steps
    n = 4096 + 4
    p0 = torch.randn(n, device="cuda", generator=g)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=5e-4, weight_decay=2e-4)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.ones(1, dtype=torch.int64, device="cuda")
    rngc = torch.zeros(2, dtype=torch.int64, device="cuda")
    for s in range(3):
        gr = torch.randn(n, device="cuda", generator=g) * (10.0 ** (s - 1))
        pt.grad = gr.clone()
        opt.step()
        ops.adam_flat(p, gr, m, v, n, 5e-4, 0.9, 0.999, 1e-8, 2e-4, step, 1.0)
        ops.advance_counters(rngc, step)
    assert step.item() == 4 and rngc[1].item() == 3
    assert relerr(p, pt.detach()) < 1e-6


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("dt", DT)
def test_pack_weights(ops, mir, dt, fused):
    from multi_modal_csi_b200 import layout as LY
    g = LY.ModelGeom(400, 30, 12)
    arena = LY.build_arena(LY.parameter_specs(g))
    plan = LY.build_pack_plan(g, arena, fused)
    params = torch.randn(arena.size, device="cuda", generator=gen(14))
    outs = []
    for o in (ops, mir):
        packed = torch.zeros(plan.size, dtype=dt, device="cuda")
        o.pack_weights(params, packed, o.make_pack_table(plan.entries, "cuda"), len(plan.entries), plan.max_elems)
        outs.append(packed.float())
    assert torch.equal(outs[0], outs[1])


def test_smooth_l1_matches_torch(ops, mir):
    g = gen(21)
    rows, cols = 37, 9
    z = torch.randn(rows, 16, device="cuda", generator=g) * 2
    y = torch.randint(0, 4, (rows, cols), device="cuda", generator=g).float()
    for o in (ops, mir):
        loss, dz = torch.zeros(1, device="cuda"), torch.zeros(rows, 16, device="cuda")
        o.smooth_l1(z, y, rows, cols, 1.0, 2.0, loss, dz)
        zz = z[:, :cols].clone().requires_grad_(True)
        ref = torch.nn.SmoothL1Loss()(zz, y)
        ref.backward()
        assert abs(loss.item() - ref.item()) < 1e-6 * max(1.0, ref.item())
        assert relerr(dz[:, :cols], 2.0 * zz.grad) < 1e-6


# ------------------------------------------------------------------------------------------------ edge cases / errors
def test_empty_batches_and_argument_errors(ops):
    """B = 0 is a no-op for every row-walking entry point (nothing launched, buffers untouched); malformed calls return a
    negative csi_status with a message in csi_last_error() -- surfaced as RuntimeError by the binding, never a crash."""
    L, d = 150, 270
    Dp = ru(d, 16)
    _, x = tokbuf(1, L, Dp, torch.float32, fill=1.0, gen=gen(20), ncols=d)
    _, y = tokbuf(1, L, Dp, torch.bfloat16)
    mean, rstd = torch.full((x.shape[0],), 7.0, device="cuda"), torch.full((x.shape[0],), 7.0, device="cuda")
    g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.layernorm_fwd(x, g, b, y, mean, rstd, 0, L, d, HALO, 1e-6)
    sums = torch.zeros(2 * Dp, dtype=torch.float64, device="cuda")
    ops.bn_stats(y, 0, L, HALO, Dp, sums)
    raw = torch.zeros(1, 3000, 270, device="cuda")
    _, left = tokbuf(1, L, Dp, torch.float32)
    _, right = tokbuf(1, d, ru(L, 16), torch.float32)
    ops.pool_dual(raw, None, None, 0, 3000, 270, None, left, right, HALO, 0, None)
    torch.cuda.synchronize()
    assert float(y.float().abs().max()) == 0.0 and float(mean.min()) == 7.0 and float(sums.abs().max()) == 0.0
    assert float(left.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="multiple of 20"):
        ops.pool_dual(raw, None, None, 1, 2999, 270, None, left, right, HALO, 0, None)
    with pytest.raises(RuntimeError, match="rng"):
        ops.pool_dual(raw, None, None, 1, 3000, 270, None, left, right, HALO, 1, None)      # augmentation without rng
    with pytest.raises(RuntimeError, match="heads"):
        z = torch.zeros(2, 7 * 16, device="cuda")
        ops.perm_ce(z, torch.zeros(2, 70, device="cuda"), 2, 7, 10, 16, 1.0, torch.zeros(1, device="cuda"), None)


def test_perm_ce_matches_mirror(ops, mir):
    """csi_perm_ce against the torch mirror on random logits with duplicated target classes.  (Permutations that tie
    exactly in real arithmetic -- two slots with the same class -- differ by summation order in fp32, so WHICH of them
    wins is implementation noise in the reference as well; loss and gradient do not depend on it.  The reference's own
    tie sample with two identical heads is pinned by the fixture test in test_gpu_model.py.)"""
    B, H, C, cp = 300, 5, 10, 16
    g = gen(21)
    z = torch.zeros(B, H * cp, device="cuda")
    v = torch.randn(B, H, C, device="cuda", generator=g)
    z.view(B, H, cp)[:, :, :C] = v
    cls = torch.randint(0, C, (B, H), device="cuda", generator=g)
    cls[::4, 1] = cls[::4, 3]                                          # identical target classes
    y = torch.nn.functional.one_hot(cls, C).float().reshape(B, H * C).contiguous()
    outs = []
    for o in (ops, mir):
        loss, dz = torch.zeros(1, device="cuda"), torch.full((B, H * cp), 3.0, device="cuda")
        best = torch.zeros(B * H, dtype=torch.int32, device="cuda")
        o.perm_ce(z, y, B, H, C, cp, 2.0, loss, dz, best)
        outs.append((loss.clone(), dz, best))
    assert abs(outs[0][0].item() - outs[1][0].item()) < 1e-5
    assert relerr(outs[0][1], outs[1][1]) < 1e-5
    # the matched heads are a permutation, and they agree with the mirror wherever the slot's class is unique in the sample
    bk, bm = outs[0][2].view(B, H).long(), outs[1][2].view(B, H).long()
    assert bool((bk.sort(1).values == torch.arange(H, device="cuda")).all())
    uniq = (cls.unsqueeze(2) == cls.unsqueeze(1)).sum(2) == 1
    assert torch.equal(bk[uniq], bm[uniq])
    assert float(outs[0][1].view(B, H, cp)[:, :, C:].abs().max()) == 0.0                  # pad columns are written as zero
