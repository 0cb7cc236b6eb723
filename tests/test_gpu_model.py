"""Whole-path GPU parity: THAT forward / backward / BCE / Adam through libcsi_that.so against (1) the golden
fixtures generated from the unmodified reference and (2) the CPU oracle run on the same seeded inputs."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# Stated tolerances (normwise relative error  ||a-b|| / ||b||):
#   fp32 mode : 1e-4 on logits and on the full gradient vector (north_star)
#   bf16 mode : 1e-2 on logits, 3e-2 on the gradient vector (SURVEY.md section 7 / BASELINE.md section 4); measured on
#               B200 at the benchmarked batch (B=256, F=270; tests/test_gpu_round2.py prints them): 2.6e-3 / 2.9e-3, the
#               reference's own bf16 autocast: 3.1e-3 / 9.3e-3.  Multi-label predictions must be identical wherever
#               |logit_ref| exceeds the measured absolute logit error (SURVEY.md section 7 "hard parts")
TOL = {"fp32": (1e-4, 1e-4), "bf16": (1e-2, 3e-2)}


def nrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def grad_err(model, ref_grads):
    num = den = 0.0
    worst = (0.0, None)
    for k, p in model.named_parameters():
        if k in ref_grads:
            r = ref_grads[k].double()
            e = (p.grad.double().cpu() - r).norm().item()
            n = r.norm().item()
            num += e * e
            den += n * n
            if n > 1e-3 * max(den, 1e-30) ** 0.5 and e / n > worst[0]:
                worst = (e / n, k)
    return (num / den) ** 0.5, worst


def build(T, F, out, mode, sd=None, seed=39):
    from multi_modal_csi_b200 import THAT
    torch.manual_seed(seed)
    m = THAT((T, F), (out,), act_dtype=mode)
    if sd is not None:
        m.load_state_dict(sd)
    m.dropout_enabled = False
    return m.to("cuda")


def synth(B, T, F, out, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, F, generator=g) * 20
    y = (torch.rand(B, out, generator=g) < 0.15).float()
    return x, y


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_small_known_answer(gold, mode):
    g = gold("that_small.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    m = build(T, F, out, mode, sd)
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    m.train()
    logits = m(x)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))(logits, y)
    loss.backward()
    tl, tg = TOL[mode]
    assert nrel(logits, torch.from_numpy(g["logits_train"])) < tl
    assert abs(loss.item() - float(g["loss"])) < tl * 10 * abs(float(g["loss"]))
    ref_grads = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("g/")}
    ge, worst = grad_err(m, ref_grads)
    assert ge < tg, (ge, worst)
    sdm = m.state_dict()
    for k in g.files:
        if k.startswith("bn/"):
            assert nrel(sdm[k[3:]].float(), torch.from_numpy(g[k]).float()) < (1e-5 if mode == "fp32" else 1e-2), k
    m.eval()
    with torch.no_grad():
        le = m(x)
    assert nrel(le, torch.from_numpy(g["logits_eval"])) < tl


@pytest.mark.parametrize("mode,F,out", [("fp32", 270, 54), ("bf16", 270, 54), ("bf16", 540, 90), ("fp32", 540, 90)])
def test_full_size_anchor_and_oracle(gold, mode, F, out):
    """Full-size [B=4, 3000, F] case: weights re-created from seed 39 (bit-identical to the reference's init),
    compared with the committed reference outputs and with the CPU oracle's full gradient."""
    from oracle import that_oracle as O
    g = gold(f"that_anchor_{F}.npz")
    T, B = 3000, 4
    m = build(T, F, out, mode)
    for k, v in m.state_dict().items():
        assert abs(v.double().sum().item() - float(g["init_sum/" + k])) <= 1e-6 * max(1.0, float(g["init_abs/" + k])), k
    sd_cpu = copy.deepcopy({k: v.cpu() for k, v in m.state_dict().items()})
    x, y = synth(B, T, F, out)
    m.train()
    logits = m(x.cuda())
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))(logits, y.cuda())
    loss.backward()
    tl, tg = TOL[mode]
    ref_logits = torch.from_numpy(g["logits_train"])
    err_abs = (logits.cpu() - ref_logits).abs().max().item()
    assert nrel(logits, ref_logits) < tl
    assert abs(loss.item() - float(g["loss"])) < 10 * tl * float(g["loss"])
    # oracle on the box's CPU, in fp64 (the fp32 reference itself is 3e-5 away from it):
    # full gradient vector
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd_cpu.items()}
    _, _, og = O.loss_and_grads(sd64, x.double(), y.double())
    ge, worst = grad_err(m, og)
    if ge >= tg:
        # The gradient is discontinuous at the LeakyReLU kinks: with ~4e6 pre-activations a few lie within fp32
        # rounding of zero, and flipping one of them moves the gradient vector by ~1e-4 (scripts/diag_parity.py:
        # at F=540 the fp32 CPU reference itself is 1.14e-4 away from fp64 on exactly one such element).  The
        # kernel path must then agree with the fp32 reference evaluation, which took the same side of the kink.
        _, _, og32 = O.loss_and_grads(copy.deepcopy(sd_cpu), x, y)
        ge32, worst32 = grad_err(m, og32)
        assert ge32 < tg, (ge, worst, ge32, worst32)
    gn = sum(p.grad.double().pow(2).sum().item() for p in m.parameters() if p.grad is not None) ** 0.5
    assert abs(gn - float(g["grad_norm"])) < tg * float(g["grad_norm"])
    for k in ("layer_output.weight", "layer_left_gaussian.var_mu", "layer_left_gaussian.var_sigma"):
        assert nrel(dict(m.named_parameters())[k].grad, torch.from_numpy(g["g/" + k])) < 5 * tg, k
    # equal multi-label predictions after the reference decision rule (utils.py:147-183,234-239)
    m.eval()
    with torch.no_grad():
        le = m(x.cuda()).cpu()
    ref_eval = torch.from_numpy(g["logits_eval"])
    assert nrel(le, ref_eval) < tl
    users = 6
    if out % users == 0:
        mine, ref = O.predict_counts(le, users), O.predict_counts(ref_eval, users)
        margin = (le - ref_eval).abs().max().item()
        p = torch.sigmoid(ref_eval.double()).reshape(B, users, -1)
        top2 = p.topk(2, dim=2).values
        decisive = ((top2[..., 0] - 0.5).abs() > margin) & ((top2[..., 0] - top2[..., 1]) > margin)
        if bool(decisive.all()):
            assert torch.equal(mine, ref)
        else:                                   # only samples whose every user decision is outside the error band
            ok = decisive.all(dim=1)
            assert torch.equal(mine[ok], ref[ok])
    assert err_abs < 1.0


def test_fused_train_step_trajectory(gold):
    """3 Adam steps (lr 5e-4, wd 2e-4) with the fused path vs the reference's torch.optim.Adam trajectory."""
    from multi_modal_csi_b200 import FusedAdam
    g = gold("that_small.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    m = build(T, F, out, "fp32", sd)
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    m.train()
    for s in range(3):
        x, y = synth(B, T, F, out, seed=1234 + s)
        loss, _ = m.fused_train_step(x.cuda(), y.cuda(), opt, augment=False)
        assert abs(loss.item() - float(g["traj_losses"][s])) < 2e-4 * max(1.0, float(g["traj_losses"][s]))
    sdm = m.state_dict()
    for k in g.files:
        if k.startswith("traj_w/") and not k.endswith(".0.bias"):
            # conv biases under a train-mode BatchNorm have a zero true gradient: the reference moves them by
            # Adam-normalised rounding noise, so they are excluded from the trajectory comparison
            assert (sdm[k[7:]].float().cpu() - torch.from_numpy(g[k]).float()).abs().max().item() < 2e-4, k


def test_autograd_path_equals_fused_path():
    """The reference-style loop (torch loss + torch.optim.Adam on the parameter views) and the fused step agree."""
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 400, 30, 12, 5
    x, y = synth(B, T, F, out)
    ma, mf = build(T, F, out, "fp32"), build(T, F, out, "fp32")
    oa = torch.optim.Adam(ma.parameters(), lr=5e-4, weight_decay=2e-4)
    of = FusedAdam(mf.parameters(), lr=5e-4, weight_decay=2e-4)
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))
    for s in range(2):
        ma.train()
        pred = ma(x.cuda())
        la = lossf(pred, y.cuda())
        oa.zero_grad()
        la.backward()
        oa.step()
        lf, _ = mf.fused_train_step(x.cuda(), y.cuda(), of, augment=False)
        assert abs(la.item() - lf.item()) < 1e-5 * max(1.0, abs(la.item()))
    for (k, a), (_, b) in zip(ma.state_dict().items(), mf.state_dict().items()):
        if k.endswith("in_proj_bias") or k.endswith(".0.bias"):
            continue    # (partly) zero true gradient: Adam normalises atomic-order rounding noise into +-lr steps
        assert (a.float() - b.float()).abs().max().item() < 2e-5, k


def test_dropout_and_augmentation_step_runs_and_is_reproducible():
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 3000, 270, 54, 8
    x, y = synth(B, T, F, out)
    losses = []
    for rep in range(2):
        m = build(T, F, out, "bf16")
        m.dropout_enabled = True
        opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
        m.train()
        ls = []
        for s in range(3):
            loss, logits = m.fused_train_step(x.cuda(), y.cuda(), opt, augment=True)
            ls.append(loss.item())
            assert torch.isfinite(logits).all()
        assert all(np.isfinite(ls))
        assert all(torch.isfinite(p).all() for p in m.parameters())
        losses.append(ls)
    # same seed -> same augmentation and dropout masks; later steps differ only by fp32 atomic summation order
    assert abs(losses[0][0] - losses[1][0]) < 1e-4 * abs(losses[0][0])
    assert all(abs(a - b) < 5e-3 * abs(a) for a, b in zip(losses[0], losses[1]))
    assert losses[0][0] != losses[0][1]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_backward_uses_the_forward_dropout_masks(mode):
    """Dropout on: the analytic gradient must equal the finite-difference directional derivative of the SAME masked
    network (masks are regenerated in backward from (seed, step, site, element), never stored)."""
    T, F, out, B = 400, 30, 12, 6
    x, y = synth(B, T, F, out)
    x, y = x.cuda(), y.cuda()
    m = build(T, F, out, "fp32")
    m.dropout_enabled = True
    m.train()
    if mode == "bf16":                      # exercise the tensor-core epilogue's mask on the analytic side only
        m.configure(act_dtype="bf16")
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))
    step0 = int(m._engine_for(B).rng[1].item())      # the Philox step the forward/backward pair below draws its masks from
    loss = lossf(m(x), y)
    loss.backward()
    g = m.flat_grads.clone()
    assert torch.isfinite(g).all()
    assert int(m._rng[1].item()) == step0 + 1       # backward closed the pair: the next forward gets new masks
    m.configure(act_dtype="fp32")           # finite differences always in fp32; same seed/step -> same masks

    def pin_masks():
        eng = m._engine_for(B)
        eng.rng.copy_(torch.tensor([m.rng_seed, step0], device="cuda"))
        eng.rng_used = False                # the next train-mode forward draws from step0 again
    gen = torch.Generator(device="cuda").manual_seed(3)
    worst = 0.0
    for trial in range(3):
        v = torch.randn(g.numel(), device="cuda", generator=gen)
        v = v / v.norm()
        analytic = float((g.double() * v.double()).sum())
        eps = 2e-3
        vals = []
        for sgn in (1.0, -1.0):
            with torch.no_grad():
                m.flat_params.add_(v, alpha=sgn * eps)
                pin_masks()
                vals.append(float(lossf(m._forward_nograd(x, True), y)))
                m.flat_params.add_(v, alpha=-sgn * eps)
        fd = (vals[0] - vals[1]) / (2 * eps)
        worst = max(worst, abs(fd - analytic) / (abs(fd) + 1e-3))
    assert worst < (3e-2 if mode == "fp32" else 8e-2), worst


def test_state_dict_roundtrip_and_eval_chunking():
    T, F, out = 400, 30, 12
    m = build(T, F, out, "fp32")
    sd = copy.deepcopy(m.state_dict())
    m2 = build(T, F, out, "fp32", seed=7)
    m2.load_state_dict({k: v.cpu() for k, v in sd.items()})
    x, _ = synth(11, T, F, out)
    m.eval(); m2.eval()
    m.configure(max_batch=4)                            # 11 samples -> chunks of 4,4,3
    with torch.no_grad():
        a, b = m(x.cuda()), m2(x.cuda())
    assert a.shape == (11, out)
    assert (a - b).abs().max().item() < 1e-5


def test_no_cpu_path():
    from multi_modal_csi_b200 import THAT
    m = THAT((400, 30), (12,))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 400, 30))


def test_train_and_run_that_end_to_end(tmp_path, monkeypatch):
    """The reference-facing entry points on the GPU: train(...) (fused path) and run_that(...) on toy CSI data."""
    import os
    os.environ["WANDB_MODE"] = "disabled"
    from multi_modal_csi_b200 import run as RUN
    from multi_modal_csi_b200.preset import preset
    rng = np.random.default_rng(0)

    def mk(n):
        x = (rng.random((n, 400, 3, 10), dtype=np.float32) * 20)
        y = np.zeros((n, 6, 9), dtype=np.int64)
        for i in range(n):
            for u in range(6):
                if rng.random() > 0.5:
                    y[i, u, rng.integers(0, 9)] = 1
        return x, y
    xtr, ytr = mk(21)
    xte, yte = mk(8)
    monkeypatch.setitem(preset["nn"], "epoch", 2)
    monkeypatch.setitem(preset["nn"], "batch_size", 4)
    res = RUN.run_that(xtr, ytr, xte, yte, var_repeat=1)
    for k in ("total_error", "perfect_prediction_percentage", "accuracy", "error_per_person", "mean_count_error",
              "counting_error_perPerson", "precision", "recall", "f1_score"):
        assert k in res
    assert np.isfinite(res["total_error"]) and len(res["error_per_person"]) == 5


def test_cuda_graph_step_equals_eager_step():
    """The graph-replayed train body (forward + BCE + backward) follows the same trajectory as eager launches."""
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 400, 30, 12, 6
    x, y = synth(B, T, F, out)
    traj = []
    for graph in (False, True):
        m = build(T, F, out, "bf16")
        m.dropout_enabled = True
        m.use_cuda_graph = graph
        opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
        m.train()
        ls = []
        for s in range(5):
            loss, logits = m.fused_train_step(x.cuda(), y.cuda(), opt, augment=True)
            ls.append(loss.item())
        traj.append(ls)
        assert (len(m._engine._graphs) == 1) == graph
    assert all(abs(a - b) < 5e-3 * max(1.0, abs(a)) for a, b in zip(*traj)), traj
    assert traj[0][0] != traj[0][1]


@pytest.mark.parametrize("graph", [False, True])
def test_two_stream_concurrency_equals_sequential(graph):
    """Left and right THAT streams issued on two CUDA streams (THATEngine._fork) == the same launches on one stream,
    eagerly and through the captured graph (dropout + augmentation on: the Philox draws do not depend on the stream)."""
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 3000, 270, 54, 8
    x, y = synth(B, T, F, out)
    res = []
    for conc in (False, True):
        m = build(T, F, out, "bf16")
        m.dropout_enabled = True
        m.use_cuda_graph = graph
        opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
        m.train()
        eng = m._engine_for(B)
        eng.concurrent = conc
        ls = []
        for s in range(4):
            loss, _ = m.fused_train_step(x.cuda(), y.cuda(), opt, augment=True)
            ls.append(loss.item())
        assert (eng._side is not None) == conc
        res.append((ls, m.flat_grads.clone(), m.flat_params.clone()))
    (l0, g0, p0), (l1, g1, p1) = res
    assert all(abs(a - b) < 5e-3 * max(1.0, abs(a)) for a, b in zip(l0, l1)), (l0, l1)
    assert nrel(p1, p0) < 1e-2


def test_backward_parts_equal_whole():
    """Data-parallel split of the backward pass (part 1 + part 2, engine.buckets) == one whole backward; after part 1
    the first bucket already holds its final gradients."""
    T, F, out, B = 400, 30, 12, 6
    x, y = synth(B, T, F, out)
    m = build(T, F, out, "fp32")
    m.train()
    eng = m._engine_for(B)
    eng.forward(x.cuda(), B, training=True, dropout=False)
    eng.loss_fwd_bwd(y.cuda(), B, 4.0)
    eng.backward(None, B, dropout=False, part=0)
    whole = eng.grads.clone()
    eng.backward(None, B, dropout=False, part=1)
    (lo1, hi1), (lo2, hi2) = eng.buckets
    assert 0 < lo1 < hi1 == eng.grads.numel() and (lo2, hi2) == (0, lo1)
    torch.cuda.synchronize()
    assert nrel(eng.grads[lo1:hi1], whole[lo1:hi1]) < 1e-5
    assert float(eng.grads[lo2:hi2].abs().max()) == 0.0
    eng.backward(None, B, dropout=False, part=2)
    assert nrel(eng.grads, whole) < 1e-5


def test_predict_counts_matches_reference_rule(gold):
    """GPU decision rule == utils.process_predictions on the committed reference fixture (and the oracle)."""
    from oracle import that_oracle as O
    g = gold("metrics.npz")
    m = build(400, 30, 54, "fp32")
    logits = torch.from_numpy(g["logits"]).cuda()
    counts = m.predict_counts(logits, users=6, threshold=0.5).cpu()
    assert np.array_equal(counts.numpy(), g["counts_pred"].astype(np.int32))
    assert torch.equal(counts.double(), O.predict_counts(torch.from_numpy(g["logits"]), 6))


def test_full_batch_properties_b256():
    """BASELINE-size batch (B=256, [3000, 270]): size-independent properties of the eval path -- sample-permutation
    equivariance and invariance to how the batch is chunked -- plus a finite train step."""
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 3000, 270, 54, 256
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(B, T, F, device="cuda", generator=g) * 20
    y = (torch.rand(B, out, device="cuda", generator=g) < 0.05).float()
    m = build(T, F, out, "bf16")
    m.configure(max_batch=B)
    m.eval()
    with torch.no_grad():
        full = m(x)
        perm = torch.randperm(B, device="cuda", generator=g)
        assert torch.equal(m(x[perm]), full[perm])                       # samples are independent in eval mode
    m.configure(max_batch=96)                                            # 256 -> chunks of 96, 96, 64
    with torch.no_grad():
        assert torch.equal(m(x), full)
    m.configure(max_batch=B)
    m.train()
    m.dropout_enabled = True
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    l0, _ = m.fused_train_step(x, y, opt, augment=True)
    l0 = l0.item()
    for _ in range(3):
        l1, _ = m.fused_train_step(x, y, opt, augment=True)
    assert np.isfinite(l0) and np.isfinite(l1.item()) and l1.item() < l0     # 4 Adam steps on one batch reduce its loss
    assert int(m.state_dict()["layer_right_encoder.0.layer_cnn.2.1.num_batches_tracked"]) == 4


def test_count_pred_sibling_head_smooth_l1(gold):
    """THAT_COUNT_PRED + SmoothL1Loss (model/that_count_pred.py, train.py:91-97): first-step logits and gradients and a
    2-step Adam trajectory of the fused path against the fixture generated from the unmodified reference."""
    from multi_modal_csi_b200 import THAT_COUNT_PRED, FusedAdam
    g = gold("that_count_pred.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    x = torch.from_numpy(g["x"]).cuda()
    yc = torch.from_numpy(g["y"]).sum(axis=1).float().cuda()              # per-activity counts (train.py:91-92)
    # autograd path, first step: logits + gradients
    torch.manual_seed(39)
    m = THAT_COUNT_PRED((T, F), [out], act_dtype="fp32")
    m.load_state_dict(sd)
    m.dropout_enabled = False
    m = m.to("cuda").train()
    pred = m(x)
    assert nrel(pred, torch.from_numpy(g["logits_train"])) < 1e-4
    torch.nn.SmoothL1Loss()(pred, yc).backward()
    ge, worst = grad_err(m, {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("g/")})
    assert ge < 1e-4, (ge, worst)
    # fused path, two steps
    mf = THAT_COUNT_PRED((T, F), [out], act_dtype="fp32")
    mf.load_state_dict(sd)
    mf.dropout_enabled = False
    mf = mf.to("cuda").train()
    opt = FusedAdam(mf.parameters(), lr=5e-4, weight_decay=0)
    for s_ in range(2):
        loss, _ = mf.fused_train_step(x, yc, opt, augment=False, loss_kind="smooth_l1")
        ref = float(g["traj_losses"][s_])
        assert abs(loss.item() - ref) < 2e-4 * max(1.0, ref)
    # weight_decay = 0 here (that_count_pred.py:397): Adam's first steps move every weight by ~lr * sign(g), so an element
    # whose true gradient is below the fp32 noise floor can land 2 * lr away -- allow a 1e-3 fraction of such elements
    sdm = mf.state_dict()
    for k in g.files:
        if k.startswith("traj_w/") and not k.endswith(".0.bias"):
            diff = (sdm[k[7:]].float().cpu() - torch.from_numpy(g[k]).float()).abs()
            assert diff.max().item() < 2.1e-3 and (diff > 2e-4).float().mean().item() < 1e-3, k


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_multi_head_sibling_perm_ce(gold, mode):
    """THAT_MULTI_HEAD + PermutationMatchingLoss (model/that_multi_head.py): logits [B,5,C], loss and gradients of the
    autograd path, and a 2-step Adam trajectory of the fused path (csi_perm_ce), against the reference fixture."""
    from multi_modal_csi_b200 import THAT_MULTI_HEAD, FusedAdam, PermutationMatchingLoss
    g = gold("that_multi_head.npz")
    T, F, C, B, H = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    tl, tg = TOL[mode]
    m = THAT_MULTI_HEAD((T, F), [C], act_dtype=mode)
    m.load_state_dict(sd)
    m.dropout_enabled = False
    m = m.to("cuda").train()
    pred = m(x)
    assert pred.shape == (B, H, C)
    assert nrel(pred, torch.from_numpy(g["logits_train"])) < tl
    loss = PermutationMatchingLoss()(pred, y)
    assert abs(loss.item() - float(g["traj_losses"][0])) < 10 * tl * float(g["traj_losses"][0])
    loss.backward()
    ge, worst = grad_err(m, {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("g/")})
    assert ge < tg, (ge, worst)
    # loss kernel alone against the reference's known answers (one sample has two identical heads: tie rule)
    lp = torch.from_numpy(g["loss_pred"]).cuda().requires_grad_(True)
    lv = PermutationMatchingLoss()(lp, torch.from_numpy(g["loss_target"]).cuda())
    lv.backward()
    assert abs(lv.item() - float(g["loss_value"])) < 1e-5
    assert nrel(lp.grad, torch.from_numpy(g["loss_grad"])) < 1e-5
    if mode != "fp32":
        return
    mf = THAT_MULTI_HEAD((T, F), [C], act_dtype="fp32")
    mf.load_state_dict(sd)
    mf.dropout_enabled = False
    mf = mf.to("cuda").train()
    opt = FusedAdam(mf.parameters(), lr=5e-4, weight_decay=0)
    for s_ in range(2):
        loss, logits = mf.fused_train_step(x, y, opt, augment=False, loss_kind="perm_ce")
        ref = float(g["traj_losses"][s_])
        assert logits.shape == (B, H, C) and abs(loss.item() - ref) < 2e-4 * max(1.0, ref)
    sdm = mf.state_dict()
    for k in g.files:                       # weight_decay 0: see test_count_pred_sibling_head_smooth_l1 for the tolerance
        if k.startswith("traj_w/") and not k.endswith(".0.bias"):
            diff = (sdm[k[7:]].float().cpu() - torch.from_numpy(g[k]).float()).abs()
            assert diff.max().item() < 2.1e-3 and (diff > 2e-4).float().mean().item() < 1e-3, k


def test_multi_head_train_loop_with_schedule(monkeypatch):
    """train(..., var_mode="multi_head") (train.py:57-63,101-102): fused step with csi_perm_ce, per-step cosine schedule
    with warm-up driving FusedAdam's learning rate, multi_head count metrics at the end of every epoch."""
    from torch.utils.data import TensorDataset
    from multi_modal_csi_b200 import THAT_MULTI_HEAD, FusedAdam, PermutationMatchingLoss
    from multi_modal_csi_b200.preset import preset
    from multi_modal_csi_b200.train import train
    monkeypatch.setenv("WANDB_MODE", "disabled")
    T, F, C, H, N = 400, 30, 10, 5, 16
    g = torch.Generator().manual_seed(3)
    x = torch.rand(N, T, F, generator=g) * 20
    y = torch.nn.functional.one_hot(torch.randint(0, C, (N, H), generator=g), C).float()
    monkeypatch.setitem(preset["nn"], "epoch", 3)
    monkeypatch.setitem(preset["nn"], "scheduler", {"type": "cosine_warmup", "num_warmup_epochs": 1, "min_lr_ratio": 0.05})
    torch.manual_seed(39)
    m = THAT_MULTI_HEAD((T, F), [C], act_dtype="bf16", max_batch=4).to("cuda")
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=0)
    w0 = m.flat_params.clone()
    sd = train(m, opt, PermutationMatchingLoss(), TensorDataset(x[:12], y[:12]), TensorDataset(x[12:], y[12:]), 0.5, 4, 3,
               torch.device("cuda"), "multi_head")
    assert set(sd.keys()) == set(m.state_dict().keys()) and len(sd) == 171
    assert not torch.equal(w0, m.flat_params) and bool(torch.isfinite(m.flat_params).all())
    # 3 epochs x 3 loader batches, the last batch of every epoch skipped (train.py:81): 6 scheduler steps, warm-up of 3
    lr = opt.param_groups[0]["lr"]
    assert 0.05e-3 <= lr < 1e-3
