"""Round-2 GPU tests: oracle parity at the BENCHMARKED shapes (BASELINE configs 1 and 2), the product batch loader
(resident / stream, dense / ragged), counters that survive an engine rebuild, Philox steps that advance without the fused
optimizer, bf16 dispatch accounting, and the tcgen05 attention kernels against the mma.sync ones."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_model import TOL, build, grad_err, nrel          # noqa: E402


def bench_batch(B, F, out, seed=1234):
    import bench
    return bench.synth_batch(B, F, out, seed)


# ------------------------------------------------------------------------------------------- parity at the bench shapes
@pytest.mark.parametrize("mode,B", [("fp32", 32), ("bf16", 32), ("bf16", 256)])
def test_oracle_parity_at_benchmarked_batch(mode, B, capsys):
    """BASELINE config 1 (B=32) and config 2 (B=256) inputs exactly as bench.py draws them (front-pad zero rows, one-hot
    activity labels): logits, loss and the full gradient vector against the fp32 CPU oracle.  BatchNorm statistics, the
    tail-wave tile split and the multi-wave grids are exercised at the measured size against the ORACLE, not the mirror.
    The measured bf16 errors are printed (pytest -s) and must stay inside TOL."""
    from oracle import that_oracle as O
    T, F, out = 3000, 270, 54
    x, y = bench_batch(B, F, out)
    m = build(T, F, out, mode)
    m.configure(max_batch=B)
    # strict: any bf16 contraction that cannot run on tcgen05 is an error (B=32: the output layer's weight gradient
    # contracts over only 32 rows, below the 64-row minimum of the tcgen05 wgrad kernel, so strict applies at B=256)
    m._engine_for(B).ops.set_strict_tc(mode == "bf16" and B >= 64)
    try:
        sd_cpu = copy.deepcopy({k: v.cpu() for k, v in m.state_dict().items()})
        m.train()
        logits = m(x.cuda())
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))(logits, y.cuda())
        loss.backward()
    finally:
        m._engine.ops.set_strict_tc(False)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_logits, ref_loss, ref_grads = O.loss_and_grads(sd_cpu, x, y)
    el = nrel(logits, ref_logits)
    ge, worst = grad_err(m, ref_grads)
    with capsys.disabled():
        print(f"\n[parity B={B} {mode}] logits rel {el:.3e}  loss {loss.item():.6f} vs {ref_loss.item():.6f}  "
              f"grad rel {ge:.3e} (worst {worst[1]} {worst[0]:.2e})")
    tl, tg = TOL[mode]
    if mode == "fp32":
        tg = 3e-4          # fp32 CPU reference vs fp32 GPU: summation order over 32 x 150 tokens (the B=4 test pins 1e-4 against fp64)
    assert el < tl and ge < tg, (el, ge, worst)
    assert abs(loss.item() - ref_loss.item()) < 10 * tl * abs(ref_loss.item())
    # equal multi-label decisions (utils.py:147-183: per user the arg-max class, kept iff sigmoid > 0.5 <=> logit > 0)
    # wherever the reference decision is outside the measured logit error
    mine, ref = logits.detach().cpu().reshape(B, 6, -1), ref_logits.reshape(B, 6, -1)
    margin = (mine - ref).abs().max().item()
    top2 = ref.topk(2, dim=2).values
    decisive = (top2[..., 0].abs() > margin) & ((top2[..., 0] - top2[..., 1]) > 2 * margin)
    assert int(decisive.sum()) > B * 6 // 2
    assert torch.equal(mine.argmax(2)[decisive], ref.argmax(2)[decisive])
    assert torch.equal((mine.max(2).values > 0)[decisive], (ref.max(2).values > 0)[decisive])
    # (the count rule itself -- fp32 sigmoid saturation ties included -- is pinned bit-exactly by
    # test_gpu_model.py::test_predict_counts_matches_reference_rule on the reference fixture)


# ------------------------------------------------------------------------------------------- loader
def _ragged(N, T, F, out, seed=5):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(T - 60, T + 1, (N,), generator=g)
    lens[0] = T
    chunks = [torch.rand(int(n), F, generator=g) * 20 for n in lens]
    y = (torch.rand(N, 6, out // 6, generator=g) < 0.1).float()
    dense = torch.zeros(N, T, F)
    for i, c in enumerate(chunks):
        dense[i, T - c.shape[0]:] = c                                   # load_data.py:66-72: FRONT zero-pad
    arena = torch.cat([c.reshape(-1) for c in chunks])
    offs = torch.zeros(N, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens[:-1] * F, 0)
    return dense, arena, offs, lens.int(), y


@pytest.mark.parametrize("mode", ["resident", "stream"])
def test_batch_source_matches_dense_batches(mode):
    """CSIBatchSource (dense TensorDataset and ragged PackedCSIDataset, both modes): the logits of every batch it feeds
    are bit-identical to feeding the front-padded dense batch directly."""
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200.loader import CSIBatchSource, PackedCSIDataset
    T, F, out, N, B = 400, 30, 54, 11, 4
    dense, arena, offs, lens, y = _ragged(N, T, F, out)
    m = build(T, F, out, "fp32")
    m.eval()
    lists = [[0, 1, 2, 3], [7, 5, 10, 2], [4, 6, 8], [9, 3, 1, 0], [2, 2, 5, 6]]
    want = []
    with torch.no_grad():
        for idx in lists:
            want.append(m(dense[idx].cuda()).clone())
    for ds in (TensorDataset(dense.reshape(N, T, 3, F // 3), y), PackedCSIDataset(arena, offs, lens, F, y, T)):
        src = CSIBatchSource(ds, "cuda", B, mode=mode)
        assert src.mode == mode and src.N == N
        got, ys, nb = [], [], []
        for batch in src.batches(lists):
            eng = m._engine_for(batch.size)
            got.append(eng.forward(batch.x, batch.size, training=False, offs=batch.offs, lens=batch.lens).clone())
            ys.append(batch.y.clone())
            nb.append(batch.h2d_bytes)
        src.close()
        for idx, a, b, yy in zip(lists, got, want, ys):
            assert torch.equal(a, b)
            assert torch.equal(yy.cpu(), y[idx])
        if mode == "resident":
            assert max(nb) < 4 * 4096                                   # only the tables and labels cross PCIe per step
        else:
            assert nb[0] >= 4 * (T - 60) * F * 4
    x0, y0 = PackedCSIDataset(arena, offs, lens, F, y, T)[3]
    assert torch.equal(x0, dense[3]) and torch.equal(y0, y[3])


@pytest.mark.parametrize("loader_mode", ["resident", "stream"])
def test_train_loop_on_packed_dataset(loader_mode, monkeypatch):
    """train() on the ragged dataset through both loader modes == train() on the padded TensorDataset: same shuffles
    (same RNG draws as DataLoader(shuffle=True)), same steps, identical final weights (dropout / augmentation are
    functions of (seed, step), not of the loader)."""
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200 import THAT, FusedAdam
    from multi_modal_csi_b200.loader import PackedCSIDataset
    from multi_modal_csi_b200.train import train
    monkeypatch.setenv("WANDB_MODE", "disabled")
    T, F, out, N, B = 400, 30, 54, 22, 4
    dense, arena, offs, lens, y = _ragged(N, T, F, out)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))
    finals = []
    for ds in (TensorDataset(dense[:16], y[:16]), PackedCSIDataset(arena[:int(offs[16])], offs[:16], lens[:16], F, y[:16], T)):
        torch.manual_seed(39)
        m = THAT((T, F), (out,), act_dtype="fp32", max_batch=B).to("cuda")
        m.rng_seed = 7
        opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
        torch.manual_seed(123)                                          # the shuffles
        sd = train(m, opt, loss, ds, TensorDataset(dense[16:], y[16:]), 0.5, B, 2, torch.device("cuda"), "baseline",
                   loader_mode=loader_mode)
        assert len(sd) == 163
        finals.append(m.flat_params.clone())
        assert int(m._opt_step.item()) == 1 + 2 * 3                     # 2 epochs x (4 batches - the skipped last one)
    # (split-K atomics make a step's gradient order-dependent in the last bits; a wrong sample or pad would be ~1e-2 away)
    assert nrel(finals[0], finals[1]) < 1e-4


# ------------------------------------------------------------------------------------------- counters / RNG
def test_engine_rebuild_keeps_adam_step_and_rng():
    """Reference call pattern THAT(x_shape, y_shape) without max_batch: train at B=8, evaluate N=20 > 8 (the engine is
    rebuilt for the larger batch), train again -- must equal the uninterrupted run bit for bit (Adam bias correction and
    the dropout / augmentation stream continue)."""
    from multi_modal_csi_b200 import THAT, FusedAdam
    T, F, out = 400, 30, 12
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(8, T, F, generator=g) * 20).cuda()
    y = (torch.rand(8, out, generator=g) < 0.2).float().cuda()
    xe = (torch.rand(20, T, F, generator=g) * 20).cuda()
    finals = []
    for interrupt in (False, True):
        torch.manual_seed(39)
        m = THAT((T, F), (out,), act_dtype="fp32").to("cuda")
        m.rng_seed = 11
        opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
        m.train()
        for _ in range(2):
            m.fused_train_step(x, y, opt, augment=True)
        if interrupt:
            e0 = m._engine
            m.eval()
            with torch.no_grad():
                m(xe)
            assert m._engine is not e0 and m._engine.B >= 20
            m.train()
        for _ in range(2):
            m.fused_train_step(x, y, opt, augment=True)
        assert int(m._opt_step.item()) == 5 and int(m._rng[1].item()) == 4
        finals.append(m.flat_params.clone())
    # a restarted Adam step (bias correction of step 1 applied at step 3) or a restarted dropout stream is ~1e-3 away;
    # run-to-run noise of the split-K atomics is ~1e-6
    assert nrel(finals[0], finals[1]) < 2e-5


def test_rng_step_advances_without_fused_adam():
    """Generic autograd path (any torch optimizer): consecutive train-mode forwards draw different dropout masks, and a
    forward / backward pair uses ONE step (backward regenerates the masks its forward drew)."""
    from multi_modal_csi_b200 import THAT
    T, F, out = 400, 30, 12
    torch.manual_seed(39)
    m = THAT((T, F), (out,), act_dtype="fp32").to("cuda")
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(4, T, F, generator=g) * 20).cuda()
    m.train()
    a = m(x).detach().clone()
    b = m(x).detach().clone()
    assert not torch.equal(a, b)                                        # two forwards in a row: new masks
    with torch.no_grad():
        c = m(x).clone()
        d = m(x).clone()
    assert not torch.equal(c, d)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    s0 = int(m._rng[1].item())
    for _ in range(3):
        opt.zero_grad()
        m(x).sum().backward()
        opt.step()
    assert int(m._rng[1].item()) == s0 + 3 + 1                          # +1: the pending no-grad forward above
    # finite-difference check that backward used the masks of its own forward: directional derivative along the gradient
    m.rng_seed = 5
    p = dict(m.named_parameters())["layer_output.bias"]
    step = int(m._rng[1].item())
    opt.zero_grad()
    out0 = m(x).sum()
    out0.backward()
    gsum = p.grad.sum().item()
    assert abs(gsum - 4 * out) < 1e-3 * 4 * out                         # d(sum logits)/d(bias_j) = B for every j
    assert int(m._rng[1].item()) == step + 1


def test_fused_adam_frozen_parameters_stay_fixed():
    from multi_modal_csi_b200 import THAT, FusedAdam
    T, F, out = 400, 30, 12
    torch.manual_seed(39)
    m = THAT((T, F), (out,), act_dtype="fp32").to("cuda")
    named = dict(m.named_parameters())
    frozen = [k for k in named if k.startswith("layer_right_")]
    for k in frozen:
        named[k].requires_grad_(False)
    with pytest.raises(ValueError):
        FusedAdam(m.parameters(), lr=1e-3, weight_decay=2e-4)
    opt = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0)
    before = {k: named[k].detach().clone() for k in named}
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(4, T, F, generator=g) * 20).cuda()
    y = (torch.rand(4, out, generator=g) < 0.2).float().cuda()
    m.train()
    for _ in range(2):
        m.fused_train_step(x, y, opt, augment=False)
    for k in named:
        if ".layer_cnn." in k and k.endswith(".0.bias"):
            continue                                                    # conv bias in front of a train-mode BatchNorm: zero gradient
        same = torch.equal(named[k].detach(), before[k])
        assert same == (k in frozen or k == "layer_left_gaussian.var_position"), k
    sd = opt.state_dict()
    assert sd["csi_flat"]["step"] == 3 and sd["csi_flat"]["exp_avg"].shape == m.flat_params.shape
    opt2 = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2._m.cpu(), sd["csi_flat"]["exp_avg"])


# ------------------------------------------------------------------------------------------- dispatch accounting
def test_bf16_dispatch_is_counted_and_strict_mode_raises():
    from multi_modal_csi_b200.ops import NativeOps
    ops = NativeOps(torch.device("cuda", 0))
    ops.dispatch_counts(reset=True)
    A = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
    W = torch.randn(32, 64, device="cuda").to(torch.bfloat16)
    Cm = torch.zeros(256, 32, dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt(A, W, Cm, 256, 32, [(0, 0, 0, 64)], None, None, 0.0, 0, None)
    assert ops.dispatch_counts() == {"tcgen05": 1, "ffma_fallback": 0, "mma_sync": 0}
    # a weight gradient that contracts over fewer than 64 rows: the tcgen05 wgrad kernel cannot take it
    G = torch.zeros(32, 64, device="cuda")
    ops.gemm_tn(A[:32, :32].contiguous(), A[:32], G, 64, 1, 32, 32, [(0, 0, 0, 64)])
    assert ops.dispatch_counts()["ffma_fallback"] == 1
    ops.set_strict_tc(True)
    try:
        with pytest.raises(RuntimeError, match="strict"):
            ops.gemm_tn(A[:32, :32].contiguous(), A[:32], G, 64, 1, 32, 32, [(0, 0, 0, 64)])
    finally:
        ops.set_strict_tc(False)
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------- tcgen05 attention
@pytest.mark.parametrize("B,L,d", [(2, 20, 30), (3, 150, 270), (2, 270, 150), (2, 150, 540), (2, 129, 160), (1, 16, 640)])
def test_attention_tcgen05_kernels(B, L, d):
    """attention_tc.cu / attention_tc_bwd.cu (S, dP, P, dS in TMEM; UTCHMMA in SASS) against SDPA autograd in fp32 on the
    bf16-rounded operands, for head widths 3/15/16/27/54/64, ragged 128-row tiles and the in_proj bias gradient."""
    from multi_modal_csi_b200.ops import NativeOps
    ops = NativeOps(torch.device("cuda", 0))
    H, HALO, GUARD = 10, 2, 16
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    Lp = L + 2 * HALO

    def mk(nsec, seed, scale=1.0):
        g = torch.Generator(device="cuda").manual_seed(seed)
        full = torch.zeros(B * Lp + 2 * GUARD, nsec * H * hp, dtype=torch.bfloat16, device="cuda")
        body = full[GUARD:GUARD + B * Lp]
        v = torch.randn(B, L, nsec * H, hd, device="cuda", generator=g) * scale
        body.view(B, Lp, nsec * H, hp)[:, HALO:HALO + L, :, :hd] = v.to(torch.bfloat16)
        return body

    def valid(body, nh):
        return body.view(B, Lp, nh, hp)[:, HALO:HALO + L, :, :hd].float()

    qkv, do = mk(3, 1), mk(1, 2)
    t = valid(qkv, 3 * H).reshape(B, L, 3, H, hd)
    q, k, v = [t[:, :, w].transpose(1, 2).clone().requires_grad_(True) for w in range(3)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    ref.backward(valid(do, H).transpose(1, 2))
    ops.set_attn_impl(1, 1)
    try:
        ops.dispatch_counts(reset=True)
        o, dqkv = mk(1, 0, 0.0), mk(3, 0, 0.0)
        lse = torch.zeros(B * H * L, device="cuda")
        dbias = torch.full((3 * d,), 0.5, device="cuda")
        ops.attn_fwd(qkv, o, lse, B, L, d, H, hp, HALO)
        ops.attn_bwd(qkv, o, do, dqkv, lse, B, L, d, H, hp, HALO, dbias)
        torch.cuda.synchronize()
        assert ops.dispatch_counts()["tcgen05"] == 2
    finally:
        ops.set_attn_impl(0, 0)
    assert nrel(valid(o, H), ref.transpose(1, 2)) < 5e-3
    gq = valid(dqkv, 3 * H).reshape(B, L, 3, H, hd)
    for w, xg in enumerate((q, k, v)):
        assert nrel(gq[:, :, w].transpose(1, 2), xg.grad) < 8e-3, w
    cs = gq.reshape(B * L, 3, H, hd).sum(0).reshape(-1) + 0.5
    assert nrel(dbias, cs) < 1e-5
    if hp > hd:
        assert float(dqkv.view(B, Lp, 3 * H, hp)[:, :, :, hd:].float().abs().max()) == 0.0


# ------------------------------------------------------------------------------------------- weight-gradient reduction modes
@pytest.mark.parametrize("M,Na,Dp,nlen,k", [(39424, 270, 272, 270, 5), (70144, 150, 160, 150, 3), (4096, 128, 272, 270, 16),
                                             (1000, 54, 288, 288, 1)])
def test_wgrad_two_stage_reduction_is_reproducible(M, Na, Dp, nlen, k):
    """csi_gemm_tn with a registered workspace (csi_gemm_tn_workspace: partial sums of the token chunks + one fixed-order
    reduce) against the fp32 product on the bf16 operands and against the atomics mode; reruns are bit-identical, and the
    call ACCUMULATES into C like the atomics mode does."""
    import ctypes as C
    from multi_modal_csi_b200 import ops as OPS
    ops = OPS.NativeOps(torch.device("cuda", 0))
    GUARD = 16
    g = torch.Generator(device="cuda").manual_seed(5)
    A = (torch.randn(M, (Na + 15) // 16 * 16, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    full = torch.randn(M + 2 * GUARD, Dp, device="cuda", generator=g).to(torch.bfloat16)
    X = full[GUARD:GUARD + M]
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j, nlen) for j in range(k)]
    ref = torch.zeros(Na, nlen, k, device="cuda")
    for j in range(k):
        ref[:, :, j] = A[:, :Na].float().t() @ full[GUARD + j - pl:GUARD + j - pl + M, :nlen].float()

    def run(init):
        out = torch.full((Na, nlen, k), init, device="cuda")
        ops.gemm_tn(A, X, out, nlen * k, k, M, Na, segs)
        torch.cuda.synchronize()
        return out

    assert OPS.TN_TWO_STAGE
    a, b = run(0.0), run(0.0)
    assert torch.equal(a, b)
    assert nrel(a, ref) < 1e-5
    assert nrel(run(0.25), ref + 0.25) < 1e-5
    # atomics mode (no workspace registered for the stream)
    st = torch.cuda.current_stream().cuda_stream
    saved = OPS._TN_WS.pop((0, st))
    ops.lib.csi_gemm_tn_workspace(C.c_void_p(st), C.c_void_p(0), C.c_longlong(0))
    OPS.TN_TWO_STAGE = False
    try:
        c = run(0.0)
    finally:
        OPS.TN_TWO_STAGE = True
        ops.lib.csi_gemm_tn_workspace(C.c_void_p(st), C.c_void_p(saved.data_ptr()), C.c_longlong(saved.numel()))
        OPS._TN_WS[(0, st)] = saved
    assert nrel(c, ref) < 1e-5 and nrel(a, c) < 1e-5


# ------------------------------------------------------------------------------------------- fused three-branch conv
@pytest.mark.parametrize("M,d,kernels", [(39424, 270, (1, 3, 5)), (70144, 150, (1, 2, 3)), (1200, 540, (1, 3, 5)), (300, 30, (1, 3, 5))])
def test_banded_gemm_is_three_convs(M, d, kernels):
    """csi_gemm_nt_banded on the stacked, tap-aligned weights of three Conv1d branches == three csi_gemm_nt calls (bit for bit:
    the same taps in the same K order per output column) == conv1d of torch on the bf16 operands."""
    from multi_modal_csi_b200.ops import NativeOps
    ops = NativeOps(torch.device("cuda", 0))
    GUARD, Dp = 16, (d + 15) // 16 * 16
    g = torch.Generator(device="cuda").manual_seed(3)
    full = torch.zeros(M + 2 * GUARD, Dp, dtype=torch.bfloat16, device="cuda")
    full[GUARD:GUARD + M, :d] = torch.randn(M, d, device="cuda", generator=g).to(torch.bfloat16)
    A = full[GUARD:GUARD + M]
    shifts = sorted({t - (k - 1) // 2 for k in kernels for t in range(k)})
    W = [(torch.randn(d, d, k, device="cuda", generator=g) / (d * k) ** 0.5).to(torch.bfloat16) for k in kernels]
    fused = torch.zeros(3 * Dp, len(shifts) * Dp, dtype=torch.bfloat16, device="cuda")
    sep = []
    for j, (k, w) in enumerate(zip(kernels, W)):
        pl = (k - 1) // 2
        s = torch.zeros(d, k * Dp, dtype=torch.bfloat16, device="cuda")
        for t in range(k):
            s[:, t * Dp:t * Dp + d] = w[:, :, t]
            c = shifts.index(t - pl)
            fused[j * Dp:j * Dp + d, c * Dp:c * Dp + d] = w[:, :, t]
        sep.append(s)
    segs = [(sh, 0, t * Dp, Dp) for t, sh in enumerate(shifts)]
    bands = []
    for sh in shifts:
        has = [j for j, k in enumerate(kernels) if -((k - 1) // 2) <= sh <= k - 1 - (k - 1) // 2]
        bands.append((has[0] * Dp, (has[-1] + 1) * Dp))
    z1 = torch.full((M, 3 * Dp), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_banded(A, fused, z1, M, 3 * Dp, segs, bands, None, None, 0.0, 0, None)
    z3 = torch.zeros(M, 3 * Dp, dtype=torch.bfloat16, device="cuda")
    for j, k in enumerate(kernels):
        pl = (k - 1) // 2
        ops.gemm_nt(A, sep[j], z3[:, j * Dp:], M, d, [(t - pl, 0, t * Dp, Dp) for t in range(k)], None, None, 0.0, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(z1, z3)
    for j, (k, w) in enumerate(zip(kernels, W)):
        pl = (k - 1) // 2
        ref = torch.zeros(M, d, device="cuda")
        for t in range(k):
            ref += full[GUARD + t - pl:GUARD + t - pl + M, :d].float() @ w[:, :, t].float().t()
        assert nrel(z1[:, j * Dp:j * Dp + d].float(), ref) < 4e-3, j
