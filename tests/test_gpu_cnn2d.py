"""CSI-as-image path (BASELINE config 4): multi_modal_csi_b200.CNN_2D through libcsi_that.so against the
fixture generated from the unmodified reference (model/cnn_2d.py) and against the CPU oracle restatement."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_model import TOL, nrel          # noqa: E402


def grad_err(model, ref_grads):
    num = den = 0.0
    worst = (0.0, None)
    for k, p in model.named_parameters():
        r = ref_grads[k].double()
        e = (p.grad.double().cpu() - r).norm().item()
        n = r.norm().item()
        num += e * e
        den += n * n
        if e / (n + 1e-12) > worst[0]:
            worst = (e / (n + 1e-12), k)
    return (num / den) ** 0.5, worst


def data(g):
    T, F, out, B = [int(v) for v in g["dims"]]
    gen = torch.Generator().manual_seed(2468)
    x = torch.rand(B, T, F, generator=gen) * 20
    y = (torch.rand(B, out, generator=gen) < 0.15).float()
    return T, F, out, B, x, y


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_im2col_and_col2im_against_unfold(dt):
    from multi_modal_csi_b200.ops import NativeOps
    ops = NativeOps(torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, H, W, C, k, s) in [(2, 40, 35, 32, 15, 3), (3, 9, 7, 64, 7, 1), (2, 61, 41, 8, 5, 2)]:
        OH, OW = (H - k) // s + 1, (W - k) // s + 1
        K = k * k * C
        Kp = (K + 15) // 16 * 16
        x = torch.randn(B, H, W, C, device="cuda", generator=g).to(dt)
        scale = torch.rand(C, device="cuda", generator=g) + 0.5
        shift = torch.randn(C, device="cuda", generator=g)
        col = torch.full((B * OH * OW, Kp), 7.0, device="cuda", dtype=dt)
        ops.im2col_bn(x, B, H, W, C, k, s, scale, shift, col, Kp)
        xn = (x.float() * scale + shift).to(dt).float()
        ref = torch.nn.functional.unfold(xn.permute(0, 3, 1, 2), k, stride=s)                # [B, C*k*k, OH*OW], (c, kh, kw) order
        ref = ref.view(B, C, k * k, OH * OW).permute(0, 3, 2, 1).reshape(B * OH * OW, K)      # -> (kh*k+kw, c)
        assert nrel(col[:, :K], ref) < (1e-6 if dt == torch.float32 else 3e-3)        # fmaf vs mul+add, then one bf16 rounding
        assert float(col[:, K:].float().abs().sum()) == 0.0
        gcol = torch.randn(B * OH * OW, Kp, device="cuda", generator=g).to(dt)
        gx = torch.full((B, H, W, C), 3.0, device="cuda")
        ops.col2im(gcol, B, H, W, C, k, s, Kp, gx)
        fold_in = gcol[:, :K].float().view(B, OH * OW, k * k, C).permute(0, 3, 2, 1).reshape(B, C * k * k, OH * OW)
        refg = torch.nn.functional.fold(fold_in, (H, W), k, stride=s).permute(0, 2, 3, 1)
        assert nrel(gx, refg) < 1e-6
    # single-channel fp32 image -> patch matrix (conv 0)
    B, H, W, k, s = 2, 300, 270, 27, 7
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    x = torch.rand(B, H, W, device="cuda", generator=g) * 20
    scale, shift = torch.tensor([0.37], device="cuda"), torch.tensor([-1.5], device="cuda")
    col = torch.full((B * OH * OW, 736), 7.0, device="cuda", dtype=dt)
    ops.im2col_bn(x, B, H, W, 1, k, s, scale, shift, col, 736)
    ref = torch.nn.functional.unfold(torch.addcmul(shift, x, scale)[:, None], k, stride=s).permute(0, 2, 1).reshape(-1, 729)
    assert nrel(col[:, :729], ref.to(dt)) < (1e-6 if dt == torch.float32 else 3e-3) and float(col[:, 729:].float().abs().sum()) == 0.0


def test_nhwc_stats_and_bn2d_backward_against_torch():
    from multi_modal_csi_b200.ops import NativeOps
    ops = NativeOps(torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(1)
    for rows, C in [(1000, 32), (333, 64), (77, 128)]:
        x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 1)
        sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
        ops.nhwc_stats(x, rows, C, sums)
        assert nrel(sums[:C], x.double().sum(0)) < 1e-6 and nrel(sums[C:], (x.double() ** 2).sum(0)) < 1e-6
    x1 = torch.rand(12345, device="cuda", generator=g)
    s1 = torch.zeros(2, dtype=torch.float64, device="cuda")
    ops.nhwc_stats(x1, 12345, 1, s1)
    assert nrel(s1, torch.stack([x1.double().sum(), (x1.double() ** 2).sum()])) < 1e-6
    # BatchNorm backward fused with the LeakyReLU backward of the producing block (dropout off), spatial-mean gradient
    B, P, C = 5, 7, 32
    z = torch.randn(B * P, C, device="cuda", generator=g)
    zr = z.clone().requires_grad_(True)
    gamma = (torch.rand(C, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = torch.zeros(C, device="cuda", requires_grad=True)
    y = torch.nn.functional.leaky_relu(zr, 0.01)
    out = torch.nn.functional.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    feat = out.view(B, P, C).mean(1)
    gfeat = torch.randn(B, C, device="cuda", generator=g)
    feat.backward(gfeat)
    yd = y.detach()
    mean, var = yd.mean(0), yd.var(0, unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    red = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    ops.bn2d_bwd_reduce(gfeat, P, 1.0 / P, yd, B * P, C, mean, invstd, red)
    gz = torch.zeros_like(z)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn2d_bwd_apply(gfeat, P, 1.0 / P, yd, z, None, 0.0, B * P, C, mean, invstd, gamma.detach(), red, gz, dg, db)
    assert nrel(gz, zr.grad) < 1e-5 and nrel(dg, gamma.grad) < 1e-5 and nrel(db, beta.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ whole model
def test_init_matches_reference_fixture(gold):
    from multi_modal_csi_b200 import CNN_2D
    g = gold("cnn2d_anchor.npz")
    T, F, out, B, x, y = data(g)
    torch.manual_seed(39)
    m = CNN_2D((T, F), (out,))
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    for k, v in sd.items():
        assert abs(v.double().sum().item() - float(g["init_sum/" + k])) <= 1e-6 * max(1.0, float(g["init_abs/" + k])), k


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cnn2d_matches_reference_fixture_and_oracle(gold, mode, capsys):
    """Train-mode logits, loss, every gradient (against the oracle's tensors and the reference's norms), running statistics
    and eval-mode logits of multi_modal_csi_b200.CNN_2D on the fixture input."""
    from multi_modal_csi_b200 import CNN_2D
    from oracle import cnn2d_oracle as C
    g = gold("cnn2d_anchor.npz")
    T, F, out, B, x, y = data(g)
    torch.manual_seed(39)
    m = CNN_2D((T, F), (out,), act_dtype=mode)
    sd_cpu = copy.deepcopy(m.state_dict())
    m.dropout_enabled = False
    m = m.to("cuda").train()
    logits = m(x.cuda())
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 6.0, device="cuda"))(logits, y.cuda())
    loss.backward()
    ref_logits, ref_loss, ref_grads = C.loss_and_grads(sd_cpu, x, y)
    # Stated tolerances.  fp32 mode: 1e-4 on logits and on the full gradient vector, against the reference fixture.
    # bf16 mode: the gradient of this network is ill-conditioned in bf16 -- every block ends in a batch-statistics
    # BatchNorm whose backward keeps only the residual orthogonal to (1, xhat), so the 2^-9 rounding of the stored
    # activations is amplified ~20x (the SAME roundings inserted into the fp32 reference algorithm,
    # cnn2d_oracle.cnn2d_forward(emulate_bf16=True), move its gradient by 9e-2..1.4e-1 and its logits by 7e-3..1e-2).
    # So bf16 parity is stated against that emulation (2e-2 logits / 5e-2 gradient: what is left is summation order and
    # rounding-boundary flips), and against the exact reference with the intrinsic bound 2e-2 / 2e-1.
    tl, tg = TOL[mode] if mode == "fp32" else (2e-2, 2e-1)
    el = nrel(logits, torch.from_numpy(g["logits_train"]))
    ge, worst = grad_err(m, ref_grads)
    with capsys.disabled():
        print(f"\n[cnn2d {mode}] logits rel {el:.3e}  loss {loss.item():.6f} vs {float(g['loss']):.6f}  grad rel {ge:.3e} (worst {worst})")
    if mode == "bf16":
        em_logits, _, em_grads = C.loss_and_grads(sd_cpu, x, y, emulate_bf16=True)
        ele, (gee, worste) = nrel(logits, em_logits), grad_err(m, em_grads)
        with capsys.disabled():
            print(f"[cnn2d bf16 vs the reference algorithm with bf16-rounded operands] logits rel {ele:.3e}  grad rel {gee:.3e} (worst {worste})")
        assert ele < 2e-2 and gee < 5e-2, (ele, gee, worste)
    assert el < tl and nrel(logits, ref_logits) < tl
    assert abs(loss.item() - float(g["loss"])) < 10 * tl * float(g["loss"])
    assert ge < tg, (ge, worst)
    if mode == "fp32":
        for k, p in m.named_parameters():
            ref = float(g["gnorm/" + k])
            assert abs(p.grad.double().norm().item() - ref) <= 5 * tg * ref + 1e-7, k
    sdm = m.state_dict()
    for k in sdm:
        if "running" in k:
            assert nrel(sdm[k].float(), torch.from_numpy(g["stat/" + k])) < (1e-5 if mode == "fp32" else 1e-2), k
        if k.endswith("num_batches_tracked"):
            assert int(sdm[k]) == 1
    m.eval()
    with torch.no_grad():
        le = m(x.cuda())
    assert nrel(le, torch.from_numpy(g["logits_eval"])) < tl


def test_cnn2d_fused_step_trajectory_and_train_loop(gold, monkeypatch):
    """fused_train_step (forward + BCE(pos_weight 6) + backward + FusedAdam(weight_decay 1e-4), cnn_2d.py:162-166) for three
    steps against the oracle with torch.optim.Adam; then train() end to end (augmentation + dropout on) through the loader."""
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200 import CNN_2D, FusedAdam
    from multi_modal_csi_b200.train import train
    from oracle import cnn2d_oracle as C
    monkeypatch.setenv("WANDB_MODE", "disabled")
    g = gold("cnn2d_anchor.npz")
    T, F, out, B, x, y = data(g)
    torch.manual_seed(39)
    m = CNN_2D((T, F), (out,), act_dtype="fp32")
    sd = copy.deepcopy(m.state_dict())
    m.dropout_enabled = False
    m = m.to("cuda").train()
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    ropt = torch.optim.Adam([params[k] for k in names], lr=1e-3, weight_decay=1e-4)
    for step in range(3):
        loss, _ = m.fused_train_step(x.cuda(), y.cuda(), opt, pos_weight=6.0, augment=False)
        work = dict(sd)
        work.update(params)
        rl = C.bce_with_logits(C.cnn2d_forward(work, x, training=True, update_stats=True), y, 6.0)
        ropt.zero_grad()
        rl.backward()
        ropt.step()
        assert abs(loss.item() - rl.item()) < 2e-4 * max(1.0, rl.item()), step
    sdm = m.state_dict()
    for k in names:
        diff = (sdm[k].float().cpu() - params[k].detach()).abs()
        assert diff.max().item() < 2.1e-3 and (diff > 2e-4).float().mean().item() < 2e-3, k       # Adam's sign-like first steps
    for k in sd:
        if "running" in k:
            assert nrel(sdm[k].float(), sd[k].float()) < 1e-4, k
    # train(): fused path through CSIBatchSource (gather + augmentation kernel), dropout on
    N = 14
    gen = torch.Generator().manual_seed(3)
    xs = torch.rand(N, T, F, generator=gen) * 20
    ys = (torch.rand(N, 6, out // 6, generator=gen) < 0.1).float()
    torch.manual_seed(39)
    m2 = CNN_2D((T, F), (out,), act_dtype="bf16", max_batch=4).to("cuda")
    opt2 = FusedAdam(m2.parameters(), lr=1e-3, weight_decay=1e-4)
    w0 = m2.flat_params.clone()
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 6.0, device="cuda"))
    out_sd = train(m2, opt2, lossf, TensorDataset(xs[:10], ys[:10]), TensorDataset(xs[10:], ys[10:]), 0.5, 4, 2, torch.device("cuda"),
                   "baseline")
    assert list(out_sd.keys()) == list(m2.state_dict().keys())
    assert bool(torch.isfinite(m2.flat_params).all()) and not torch.equal(w0, m2.flat_params)
    assert int(m2._opt_step.item()) == 1 + 2 * 2 and int(m2.state_dict()["layer_norm_2.num_batches_tracked"]) == 4


def test_cnn2d_dropout_masks_forward_and_backward_agree_with_the_oracle(gold):
    """Dropout(0.2) on: the keep bits the forward kernel stored are read back and handed to the oracle as ITS dropout
    masks; logits and every gradient must then match to fp32 accuracy -- i.e. forward and backward used the same masks, with
    the 1/(1-p) scaling, and the keep rate is 0.8."""
    from multi_modal_csi_b200 import CNN_2D
    from oracle import cnn2d_oracle as C
    g = gold("cnn2d_anchor.npz")
    T, F, out, B, x, y = data(g)
    torch.manual_seed(39)
    m = CNN_2D((T, F), (out,), act_dtype="fp32")
    sd_cpu = copy.deepcopy(m.state_dict())
    m = m.to("cuda").train()
    logits = m(x.cuda())
    torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 6.0, device="cuda"))(logits, y.cuda()).backward()
    eng, geo = m._engine, m.geom
    masks, kept = [], []
    for i in range(3):
        M, co = geo.rows(i + 1, B), geo.C[i + 1]
        bits = eng.L[i]["mask"][:M * co // 8].cpu().to(torch.int64)
        keep = ((bits[:, None] >> torch.arange(8)) & 1).reshape(B, geo.H[i + 1], geo.W[i + 1], co).permute(0, 3, 1, 2).float()
        masks.append(keep)
        kept.append(keep.mean().item())
    assert abs(kept[0] - 0.8) < 0.01 and abs(kept[1] - 0.8) < 0.02, kept
    it = iter(masks)
    names = [k for k, v in sd_cpu.items() if v.is_floating_point() and "running" not in k]
    leaves = {k: sd_cpu[k].clone().requires_grad_(True) for k in names}
    work = dict(sd_cpu)
    work.update(leaves)
    ref_logits = C.cnn2d_forward(work, x, training=True, update_stats=False, drop=lambda t, p: t * next(it) / (1.0 - p))
    ref_loss = C.bce_with_logits(ref_logits, y, 6.0)
    ref_grads = dict(zip(names, torch.autograd.grad(ref_loss, [leaves[k] for k in names])))
    ge, worst = grad_err(m, ref_grads)
    assert nrel(logits, ref_logits) < 1e-4 and ge < 1e-4, (nrel(logits, ref_logits), ge, worst)
    # a second forward/backward pair draws new masks
    m(x.cuda()).sum().backward()
    assert not torch.equal(bits, eng.L[2]["mask"][:bits.numel()].cpu().to(torch.int64))
