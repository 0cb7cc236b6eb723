"""CPU suite: oracle vs golden vectors (and vs the live reference when /root/reference exists), host logic
(labels, metrics, engine sequencing with the torch mirror, optimizer, loader), and the C-ABI symbol table."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from mirror_ops import MirrorOps
from oracle import that_oracle as O
from oracle.ref_import import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nrel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


# ------------------------------------------------------------------------------------------------ oracle pinned
def test_oracle_matches_golden_small(gold):
    g = gold("that_small.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w/")}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits, loss, grads = O.loss_and_grads(sd, x, y)
    assert nrel(logits, g["logits_train"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    num = sum((grads[k] - torch.from_numpy(g["g/" + k])).pow(2).sum().item() for k in grads)
    den = sum(torch.from_numpy(g["g/" + k]).pow(2).sum().item() for k in grads)
    assert (num / den) ** 0.5 < 1e-5
    for k in g.files:
        if k.startswith("bn/"):
            assert nrel(sd[k[3:]].float(), torch.from_numpy(g[k]).float()) < 1e-6, k
    assert nrel(O.that_forward(sd, x, training=False), g["logits_eval"]) < 1e-5


def test_oracle_adam_trajectory_matches_golden(gold):
    g = gold("that_small.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w/")}
    opt = {}
    for s in range(3):
        gg = torch.Generator().manual_seed(1234 + s)
        x = torch.rand(B, T, F, generator=gg) * 20
        y = (torch.rand(B, out, generator=gg) < 0.15).float()
        _, loss = O.train_step(sd, opt, x, y)
        assert abs(loss.item() - float(g["traj_losses"][s])) < 2e-5 * max(1, float(g["traj_losses"][s]))


@pytest.mark.parametrize("F,out", [(270, 54)])
def test_oracle_matches_full_size_anchor(gold, F, out):
    from multi_modal_csi_b200 import THAT
    g = gold(f"that_anchor_{F}.npz")
    torch.manual_seed(39)
    m = THAT((3000, F), (out,))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():                       # our initialisation == the reference's, tensor by tensor
        assert abs(v.double().sum().item() - float(g["init_sum/" + k])) <= 1e-6 * max(1.0, float(g["init_abs/" + k])), k
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(4, 3000, F, generator=gen) * 20
    y = (torch.rand(4, out, generator=gen) < 0.15).float()
    logits, loss, grads = O.loss_and_grads(sd, x, y)
    assert nrel(logits, g["logits_train"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    gn = sum(v.double().pow(2).sum().item() for v in grads.values()) ** 0.5
    assert abs(gn - float(g["grad_norm"])) < 1e-4 * float(g["grad_norm"])


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference checkout not present")
def test_oracle_and_init_match_live_reference():
    from oracle.ref_import import load_reference
    from multi_modal_csi_b200 import THAT
    ns = load_reference()
    T, F, out, B = 600, 40, 18, 3
    torch.manual_seed(41)
    ref = ns.that.THAT((T, F), (out,))
    torch.manual_seed(41)
    mine = THAT((T, F), (out,))
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    assert all(torch.equal(sr[k], sm[k]) for k in sr)
    for mod in ref.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(B, T, F, generator=gen) * 20
    y = (torch.rand(B, out, generator=gen) < 0.2).float()
    sd = copy.deepcopy(sr)
    ref.train()
    lo = ref(x)
    l = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * out))(lo, y)
    l.backward()
    ol, oloss, og = O.loss_and_grads(sd, x, y)
    assert nrel(ol, lo.detach()) < 1e-5 and abs(oloss.item() - l.item()) < 1e-5
    named = dict(ref.named_parameters())
    num = sum((og[k] - named[k].grad).pow(2).sum().item() for k in og)
    den = sum(named[k].grad.pow(2).sum().item() for k in og)
    assert (num / den) ** 0.5 < 1e-5


def test_oracle_front_pad_and_prediction_rule(gold):
    a = torch.arange(12.0).reshape(3, 4)
    p = O.front_pad(a, 5)
    assert p.shape == (5, 4) and float(p[:2].abs().sum()) == 0 and torch.equal(p[2:], a)
    with pytest.raises(ValueError):
        O.front_pad(a, 2)
    g = gold("metrics.npz")
    counts = O.predict_counts(torch.from_numpy(g["logits"]), 6)
    assert np.array_equal(counts.numpy(), g["counts_pred"])


# ------------------------------------------------------------------------------------------------ host logic
def test_metrics_match_reference_fixture(gold):
    from multi_modal_csi_b200.utils import performance_metrics, NumpyEncoder
    import json
    g = gold("metrics.npz")
    res = performance_metrics(g["y_true"], g["logits"], var_mode="baseline", var_threshold=0.5)
    for k, v in res.items():
        assert np.allclose(np.asarray(v, dtype=np.float64), g["m/" + k], rtol=1e-9, atol=1e-12, equal_nan=True), k
    # train.py:105-109 quirk: last-batch metrics use int-truncated logits
    resq = performance_metrics(g["y_true"].reshape(64, -1).astype(int), g["logits"].astype(int), var_mode="baseline")
    for k, v in resq.items():
        assert np.allclose(np.asarray(v, dtype=np.float64), g["mq/" + k], rtol=1e-9, atol=1e-12, equal_nan=True), k
    json.dumps(res, cls=NumpyEncoder)
    with pytest.raises(ValueError):
        performance_metrics(g["y_true"], g["logits"], var_mode="multi_head")       # needs [B, heads, classes] arrays
    with pytest.raises(ValueError):
        performance_metrics(g["y_true"], g["logits"], var_mode="no_such_mode")


def test_label_encoders_match_reference_fixture(gold):
    from multi_modal_csi_b200 import load_data as LD
    g = gold("labels.npz")
    csv = os.path.join(ROOT, "tests", "golden", "annotation_excerpt.csv")
    sel = LD.load_data_y(csv, ["classroom", "empty_room"], ["2.4"], ["0", "1", "3", "5"])
    assert sel["label"].to_list() == [str(s) for s in g["labels"]]
    assert np.array_equal(LD.encode_data_y(sel, "identity"), g["identity"])
    assert np.array_equal(LD.encode_data_y(sel, "activity"), g["activity"])
    assert np.array_equal(LD.encode_data_y(sel, "location"), g["location"])
    allsel = LD.load_data_y(csv)
    assert len(allsel) == int(g["n_all"])
    assert np.array_equal(LD.encode_data_y(allsel, "activity"), g["activity_all"])


def test_loader_front_pads_and_packs(tmp_path):
    from multi_modal_csi_b200 import load_data as LD
    from multi_modal_csi_b200.preset import preset
    rng = np.random.default_rng(0)
    lens = [3000, 2871, 17]
    for i, t in enumerate(lens):
        np.save(tmp_path / f"s{i}.npy", rng.random((t, 3, 3, 30), dtype=np.float32))
    x = LD.load_data_x(str(tmp_path), ["s0", "s1", "s2"])
    assert x.shape == (3, preset["data"]["length"], 3, 3, 30) and x.dtype == np.float32
    assert float(np.abs(x[1, :3000 - 2871]).sum()) == 0 and float(np.abs(x[2, :3000 - 17]).sum()) == 0
    arena, offs, ln, F = LD.load_data_x_packed(str(tmp_path), ["s0", "s1", "s2"])
    assert F == 270 and list(ln) == lens and arena.size == sum(lens) * 270
    for i, t in enumerate(lens):                                   # pack + front pad == reference padded array
        assert np.array_equal(arena[offs[i]:offs[i] + t * F].reshape(t, F), x[i, 3000 - t:].reshape(t, F))
    np.save(tmp_path / "long.npy", np.zeros((3001, 3, 3, 30), np.float32))
    with pytest.raises(ValueError):
        LD.load_data_x(str(tmp_path), ["long"])


def _mirror_model(T, F, out, sd=None):
    from multi_modal_csi_b200 import THAT
    m = THAT((T, F), (out,), act_dtype="fp32")
    m._ops_override = MirrorOps()
    m.dropout_enabled = False
    if sd is not None:
        m.load_state_dict(sd)
    return m


def test_engine_sequencing_matches_golden_with_mirror_kernels(gold):
    """Host-side composition (token layouts, shifted-GEMM convolutions, hand-derived backward) on CPU."""
    g = gold("that_small.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    m = _mirror_model(T, F, out, sd)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    m.train()
    logits = m(x)
    assert nrel(logits, g["logits_train"]) < 1e-5
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * out))(logits, y)
    loss.backward()
    num = den = 0.0
    for k, p in m.named_parameters():
        if "g/" + k in g.files:
            r = torch.from_numpy(g["g/" + k])
            num += (p.grad - r).pow(2).sum().item()
            den += r.pow(2).sum().item()
    assert (num / den) ** 0.5 < 1e-5
    m.eval()
    with torch.no_grad():
        assert nrel(m(x), g["logits_eval"]) < 1e-5


def test_ragged_input_equals_front_padded_dense():
    T, F, out, B = 400, 30, 12, 3
    torch.manual_seed(1)
    m = _mirror_model(T, F, out)
    lens = torch.tensor([400, 333, 20], dtype=torch.int32)
    offs = torch.zeros(B, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens[:-1].long() * F, 0)
    arena = torch.rand(int((lens.long() * F).sum())) * 20
    dense = torch.stack([O.front_pad(arena[offs[i]:offs[i] + int(lens[i]) * F].view(int(lens[i]), F), T) for i in range(B)])
    eng = m._engine_for(B)
    eng.repack()
    a = eng.forward(dense, B, training=False).clone()
    b = eng.forward(arena, B, training=False, offs=offs, lens=lens).clone()
    assert torch.equal(a, b)


def test_fused_adam_is_coupled_l2_and_skips_frozen():
    from multi_modal_csi_b200 import FusedAdam
    T, F, out, B = 400, 30, 12, 2
    torch.manual_seed(3)
    m = _mirror_model(T, F, out)
    ref = copy.deepcopy({k: v.clone() for k, v in m.state_dict().items()})
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    x = torch.rand(B, T, F) * 20
    y = (torch.rand(B, out) < 0.2).float()
    m.train()
    opt_state = {}
    for _ in range(2):
        m.fused_train_step(x, y, opt, augment=False)
        O.train_step(ref, opt_state, x, y)
    sd = m.state_dict()
    assert torch.equal(sd["layer_left_gaussian.var_position"], ref["layer_left_gaussian.var_position"])
    for k in sd:
        if k.endswith(".0.bias") or k.endswith("in_proj_bias"):
            continue
        assert (sd[k].float() - ref[k].float()).abs().max().item() < 5e-5, k
    with pytest.raises(ValueError):
        FusedAdam([torch.nn.Parameter(torch.zeros(3))])


def test_module_contract():
    from multi_modal_csi_b200 import THAT
    m = THAT((3000, 270), (54,))
    sd = m.state_dict()
    assert len(sd) == 163 and sum(p.numel() for p in m.parameters()) == 4901504
    assert len(list(m.parameters())) == 118 and sum(p.requires_grad for p in m.parameters()) == 117
    assert m.layer_left_encoder[3].layer_cnn[2][0].weight.shape == (270, 270, 5)
    assert m.layer_right_encoder[0].layer_attention.out_proj.weight.shape == (150, 150)
    m2 = THAT((3000, 270), (54,))
    m2.load_state_dict(copy.deepcopy(sd))
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    # parameters alias the flat arena
    m.layer_output.bias.data.fill_(3.0)
    off = m.arena.offsets["layer_output.bias"]
    assert float(m.flat_params[off]) == 3.0
    for bad in ((3001, 270), (3000, 275), (200, 270)):
        with pytest.raises(ValueError):
            THAT(bad, (54,))
    with pytest.raises(RuntimeError):                       # no CPU path in the product
        m(torch.zeros(1, 3000, 270))


def test_count_pred_sibling_matches_reference_fixture(gold):
    """THAT_COUNT_PRED (model/that_count_pred.py): identical initial weights under the reference seed, and the
    count-mode metrics of utils.py:229-233, both against fixtures generated from the unmodified reference."""
    from multi_modal_csi_b200 import THAT_COUNT_PRED
    from multi_modal_csi_b200.utils import performance_metrics
    g = gold("that_count_pred.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    torch.manual_seed(39)
    m = THAT_COUNT_PRED((T, F), [out])
    sd = m.state_dict()
    keys = [k[2:] for k in g.files if k.startswith("w/")]
    assert sorted(keys) == sorted(sd.keys())
    for k in keys:
        assert torch.equal(sd[k].cpu(), torch.from_numpy(g["w/" + k])), k
    res = performance_metrics(g["m_y_true"], g["m_y_pred"], var_mode="count_classification")
    for k, v in res.items():
        assert np.allclose(np.asarray(v, dtype=np.float64), g["m/" + k], rtol=1e-9, atol=1e-12, equal_nan=True), k


def test_multi_head_sibling_matches_reference_fixture(gold):
    """Five-head THAT + PermutationMatchingLoss (model/that_multi_head.py:180-342) against the fixture generated from the
    unmodified reference: initial weights under the reference seed (171 keys, heads registered first), the oracle's
    forward and loss, the loss-only known answers (ties included), and the engine's launch sequence through the mirror
    kernels (logits, loss, every gradient)."""
    from mirror_ops import MirrorOps
    from oracle import that_oracle as O
    from multi_modal_csi_b200 import THAT_MULTI_HEAD
    from multi_modal_csi_b200.utils import performance_metrics
    g = gold("that_multi_head.npz")
    T, F, C, B, H = [int(v) for v in g["dims"]]
    torch.manual_seed(39)
    m = THAT_MULTI_HEAD((T, F), [C], act_dtype="fp32")
    sd = m.state_dict()
    keys = [k[2:] for k in g.files if k.startswith("w/")]
    assert list(sd.keys()) == keys and len(keys) == 171                  # same keys in the same (registration) order
    for k in keys:
        assert torch.equal(sd[k].cpu(), torch.from_numpy(g["w/" + k])), k
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    # oracle restatement
    sd_o = {k: v.clone() for k, v in sd.items()}
    pred = O.that_forward(sd_o, x, training=True)
    assert pred.shape == (B, H, C) and torch.allclose(pred, torch.from_numpy(g["logits_train"]), rtol=1e-4, atol=1e-5)
    lv, _ = O.permutation_matching_loss(pred, y)
    assert abs(lv.item() - float(g["traj_losses"][0])) < 1e-5
    lp = torch.from_numpy(g["loss_pred"]).requires_grad_(True)
    lk, best = O.permutation_matching_loss(lp, torch.from_numpy(g["loss_target"]))
    lk.backward()
    assert abs(lk.item() - float(g["loss_value"])) < 1e-6 and torch.allclose(lp.grad, torch.from_numpy(g["loss_grad"]), atol=1e-7)
    assert sorted(best[5].tolist()) == list(range(H))                    # the sample with two identical heads
    # engine sequencing + the mirror of csi_perm_ce
    m._ops_override = MirrorOps()
    m.dropout_enabled = False
    m.train()
    eng = m._engine_for(B)
    logits = eng.forward(x, B, training=True, dropout=False)
    assert torch.allclose(logits, torch.from_numpy(g["logits_train"]), rtol=1e-4, atol=1e-5)
    eng.loss_kind = "perm_ce"
    loss = eng.loss_fwd_bwd(y.reshape(B, -1).float(), B)
    assert abs(loss.item() - float(g["traj_losses"][0])) < 1e-5
    eng.backward(None, B, dropout=False)
    m._attach_grads()
    for k, p in m.named_parameters():
        if "g/" + k in g.files and not k.endswith(".0.bias"):            # conv biases feed a train-mode BatchNorm: exactly 0 here
            r = torch.from_numpy(g["g/" + k])
            assert ((p.grad - r).norm() / (r.norm() + 1e-12)).item() < 2e-4, k
    # count metrics of the multi_head mode (utils.py:220-228 as intended: 3-D predictions, last class dropped)
    res = performance_metrics(g["y"], g["logits_train"], var_mode="multi_head")
    cnt_pred = np.eye(C)[g["logits_train"].argmax(-1)].sum(1)[:, :-1]
    assert np.isclose(res["total_error"], np.abs(g["y"].sum(1)[:, :-1] - cnt_pred).sum() / B)
    res4 = performance_metrics(g["y"], g["logits_train"][None], var_mode="multi_head")      # reference-style stacked input
    assert np.isclose(res4["total_error"], res["total_error"])
    # label transform of run_main.py:39-40 (utils.py:272-287): [6 users, 9 activities] -> [5 slots, 10 classes]
    from multi_modal_csi_b200.utils import reduce_dataset
    lab = np.zeros((2, 6, 9))
    lab[0, 0, 3] = lab[0, 2, 5] = 1
    red = reduce_dataset(lab)
    assert red.shape == (2, 5, 10) and red[0].argmax(-1).tolist() == [3, 5, 9, 9, 9] and red[1].argmax(-1).tolist() == [9] * 5
    assert np.all(red.sum(-1) == 1)


def test_cnn2d_oracle_matches_reference_fixture(gold):
    """Groundwork for the CSI-as-image path: the restatement of model/cnn_2d.py (CNN_2D) reproduces the
    reference's initial weights under its seed (same RNG draw order), logits, loss, gradient norms, running statistics
    and eval-mode logits.  No CUDA path is built on it yet."""
    from oracle import cnn2d_oracle as C
    g = gold("cnn2d_anchor.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    torch.manual_seed(39)
    sd = C.cnn2d_init(out)
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    for k, v in sd.items():
        assert abs(v.double().sum().item() - float(g["init_sum/" + k])) <= 1e-6 * max(1.0, float(g["init_abs/" + k])), k
    gen = torch.Generator().manual_seed(2468)
    x = torch.rand(B, T, F, generator=gen) * 20
    y = (torch.rand(B, out, generator=gen) < 0.15).float()
    logits, loss, grads = C.loss_and_grads(sd, x, y)
    assert torch.allclose(logits, torch.from_numpy(g["logits_train"]), rtol=1e-4, atol=2e-5)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    for k, gr in grads.items():
        ref = float(g["gnorm/" + k])
        assert abs(gr.double().norm().item() - ref) <= 2e-4 * ref + 1e-9, k
    C.cnn2d_forward(sd, x, training=True, update_stats=True)
    for k in sd:
        if "running" in k:
            assert np.allclose(sd[k].numpy(), g["stat/" + k], rtol=1e-4, atol=1e-6), k
    assert torch.allclose(C.cnn2d_forward(sd, x, training=False), torch.from_numpy(g["logits_eval"]), rtol=1e-4, atol=2e-5)


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from multi_modal_csi_b200 import ops
    hdr = open(os.path.join(ROOT, "include", "csi_that.h")).read()
    declared = set(re.findall(r"\b(csi_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ops.load_library()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(ops.EXPORTS) <= declared
    assert lib.csi_abi_version() == ops.ABI_VERSION
    assert ctypes.sizeof(ops.PackEntry) == 64 and ctypes.sizeof(ops.Seg) == 16 and ctypes.sizeof(ops.Ptr3) == 24
    assert ctypes.sizeof(ops.Grp) == 8


# ------------------------------------------------------------------------------------------------ train loop
def _toy_sets(n_train=10, n_valid=6, T=400, F=30, seed=0):
    g = torch.Generator().manual_seed(seed)
    from torch.utils.data import TensorDataset

    def mk(n):
        x = torch.rand(n, T, F, generator=g) * 20
        y = torch.zeros(n, 6, 9, dtype=torch.int64)
        for i in range(n):
            for u in range(6):
                if torch.rand((), generator=g) > 0.5:
                    y[i, u, int(torch.randint(0, 9, (), generator=g))] = 1
        return TensorDataset(x, y)
    return mk(n_train), mk(n_valid)


def test_train_loop_semantics_on_cpu_with_mirror(monkeypatch):
    """Reference loop semantics (train.py:75-176): last batch skipped, eval on the whole validation set, best weights
    returned as a state_dict with the reference's keys."""
    from multi_modal_csi_b200 import train as TR
    os.environ["WANDB_MODE"] = "disabled"
    tr, va = _toy_sets()
    torch.manual_seed(39)
    m = _mirror_model(400, 30, 54)
    calls = {"train": 0, "eval": 0}
    orig = m.forward

    def counting(x):
        calls["train" if m.training else "eval"] += 1
        return orig(x)
    monkeypatch.setattr(m, "forward", counting)
    monkeypatch.setattr(TR, "apply_augmentation", lambda x: x)
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * 54))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    best = TR.train(m, opt, loss, tr, va, 0.5, 4, 2, torch.device("cpu"), "baseline", patience=150)
    assert calls == {"train": 2 * 2, "eval": 2}            # 3 batches per epoch, the last one skipped
    assert list(best.keys()) == list(before.keys())
    assert any(not torch.equal(best[k], before[k]) for k in best if k.endswith("weight"))
    assert int(m.state_dict()["layer_left_encoder.0.layer_cnn.0.1.num_batches_tracked"]) == 4


def test_gradient_buckets_cover_the_arena_and_backward_parts_compose():
    """THATEngine.buckets: [split, n) is final after backward part 1, [0, split) after part 2; split = first parameter of
    left encoder 1; part 1 + part 2 == the whole backward (mirror kernels on CPU)."""
    from mirror_ops import MirrorOps
    from multi_modal_csi_b200 import THAT
    T, F, out, B = 400, 30, 12, 3
    torch.manual_seed(39)
    m = THAT((T, F), (out,), act_dtype="fp32")
    m._ops_override = MirrorOps()
    m.dropout_enabled = False
    m.train()
    eng = m._engine_for(B)
    (lo1, hi1), (lo2, hi2) = eng.buckets
    assert (lo2, hi1) == (0, m.flat_grads.numel()) and hi2 == lo1
    early = [k for k, off in m.arena.offsets.items() if off < lo1]
    assert early and all(k.startswith(("layer_left_gaussian.", "layer_left_encoder.0.")) for k in early)
    assert m.arena.offsets["layer_left_encoder.1.layer_norm_0.weight"] >= lo1
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, T, F, generator=g) * 20
    y = (torch.rand(B, out, generator=g) < 0.2).float()
    eng.forward(x, B, training=True, dropout=False)
    eng.loss_fwd_bwd(y, B, 4.0)
    eng.backward(None, B, dropout=False, part=0)
    whole = eng.grads.clone()
    eng.backward(None, B, dropout=False, part=1)
    assert torch.allclose(eng.grads[lo1:hi1], whole[lo1:hi1], rtol=1e-5, atol=1e-8)
    assert float(eng.grads[lo2:hi2].abs().max()) == 0.0
    eng.backward(None, B, dropout=False, part=2)
    assert torch.allclose(eng.grads, whole, rtol=1e-5, atol=1e-8)


def test_bench_family_table_and_roofline_schema():
    """bench.py's per-family table: tensor families against the bf16 peak, the others against the HBM peak."""
    import bench
    table = {"gemm_nt": (2.0, 48, {"flops": 2.0e12, "bytes": 1.0e9}), "layernorm_fwd": (0.25, 12, {"flops": 0, "bytes": 1.0e9}),
             "bn_finalize": (0.04, 5, {"flops": 0, "bytes": 0})}
    peaks = {"hbm_gbs": 6547.5, "bf16_tflops": 1595.7, "bf16_tflops_sustained": 1340.8}
    fam = bench.family_table(table, peaks)
    assert list(fam) == ["gemm_nt", "layernorm_fwd", "bn_finalize"]                    # sorted by device time
    assert fam["gemm_nt"]["bound"] == "tensor" and abs(fam["gemm_nt"]["achieved"] - 1000.0) < 1e-6
    assert abs(fam["gemm_nt"]["frac"] - round(1000.0 / 1595.7, 3)) < 1e-9          # burst cuBLAS peak
    assert fam["layernorm_fwd"]["bound"] == "hbm" and abs(fam["layernorm_fwd"]["achieved"] - 4000.0) < 1e-6
    assert "bound" not in fam["bn_finalize"] and fam["bn_finalize"]["launches"] == 5
    r = bench.roofline("gemm_nt", table, 256, peaks)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - 1000.0 / 1595.7) < 1e-9
    assert abs(r["frac_sustained"] - 1000.0 / 1340.8) < 1e-9 and r["peak"] == 1595.7
    assert set(r) >= {"kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "launches", "ms_per_launch"}


def test_cosine_schedule_with_warmup_matches_reference_formula():
    """train.py:26-33: linear warm-up, cosine decay, floor at min_lr_ratio -- checked against the closed form and, in the
    build container, against the reference's own get_cosine_schedule_with_warmup."""
    import math
    from multi_modal_csi_b200.train import cosine_schedule_with_warmup
    lr0, warm, total, floor = 2e-3, 5, 40, 0.05
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=lr0)
    sch = cosine_schedule_with_warmup(opt, warm, total, floor)
    ours = []
    for _ in range(total + 3):
        ours.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()

    def closed(s):
        if s < warm:
            return lr0 * s / warm
        return lr0 * max(floor, 0.5 * (1.0 + math.cos(math.pi * (s - warm) / (total - warm))))
    assert all(abs(a - closed(s)) < 1e-12 for s, a in enumerate(ours))
    assert ours[0] == 0.0 and abs(ours[warm] - lr0) < 1e-15 and abs(ours[-1] - lr0 * floor) < 1e-15
    from oracle.ref_import import reference_available, load_reference
    if reference_available():
        ref_train = load_reference().train
        opt2 = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr0)
        sch2 = ref_train.get_cosine_schedule_with_warmup(opt2, warm, total, floor)
        for s in range(total + 3):
            assert abs(opt2.param_groups[0]["lr"] - ours[s]) < 1e-15, s
            opt2.step()
            sch2.step()


def test_counters_live_on_the_model_and_survive_an_engine_rebuild():
    """the Adam step and the Philox {seed, step} belong to the model, not to the engine: an engine rebuilt for
    a larger batch (epoch-0 evaluation of train()) continues them, and the Philox step advances once per
    forward/backward pair whichever optimizer follows."""
    from multi_modal_csi_b200 import FusedAdam
    T, F, out = 400, 30, 12
    torch.manual_seed(3)
    m = _mirror_model(T, F, out)
    m.rng_seed = 77
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    x, y = torch.rand(2, T, F) * 20, (torch.rand(2, out) < 0.2).float()
    m.train()
    for _ in range(2):
        m.fused_train_step(x, y, opt, augment=False)
    e0 = m._engine
    assert m._rng.tolist() == [77, 2] and int(m._opt_step) == 3 and e0.rng is m._rng
    m.eval()
    with torch.no_grad():
        m(torch.rand(5, T, F))                                   # N=5 > engine batch 2: rebuild
    assert m._engine is not e0 and m._engine.rng is m._rng and m._engine.opt_step is m._opt_step
    m.train()
    m.fused_train_step(x, y, opt, augment=False)
    assert m._rng.tolist() == [77, 3] and int(m._opt_step) == 4
    # autograd path with a torch optimizer: one Philox step per forward/backward pair, none from the optimizer
    sgd = torch.optim.SGD(m.parameters(), lr=0.0)
    for _ in range(2):
        sgd.zero_grad()
        m(x).sum().backward()
        sgd.step()
    assert m._rng.tolist() == [77, 5] and int(m._opt_step) == 4
    m(x)
    m(x)                                                         # two train forwards without backward: masks are not reused
    assert int(m._rng[1]) == 6
    m.rng_seed = 5
    assert m._rng.tolist() == [5, 6]


def test_fused_adam_keeps_omitted_parameters_fixed():
    from multi_modal_csi_b200 import FusedAdam
    T, F, out = 400, 30, 12
    torch.manual_seed(3)
    m = _mirror_model(T, F, out)
    named = dict(m.named_parameters())
    frozen = [k for k in named if k.startswith("layer_right_")]
    for k in frozen:
        named[k].requires_grad_(False)
    with pytest.raises(ValueError):
        FusedAdam(m.parameters(), lr=1e-3, weight_decay=2e-4)    # coupled L2 would move the frozen weights
    opt = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0)
    before = {k: v.detach().clone() for k, v in named.items()}
    x, y = torch.rand(2, T, F) * 20, (torch.rand(2, out) < 0.2).float()
    m.train()
    for _ in range(2):
        m.fused_train_step(x, y, opt, augment=False)
    for k in named:
        if ".layer_cnn." in k and k.endswith(".0.bias"):
            continue                                                 # conv bias in front of a train-mode BatchNorm: zero gradient
        assert torch.equal(named[k].detach(), before[k]) == (k in frozen or k == "layer_left_gaussian.var_position"), k
    sd = opt.state_dict()
    assert sd["csi_flat"]["step"] == 3
    opt2 = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2._m, sd["csi_flat"]["exp_avg"])


def test_epoch_index_batches_match_dataloader_and_packed_dataset_pads_in_front():
    from torch.utils.data import DataLoader, TensorDataset
    from multi_modal_csi_b200.loader import PackedCSIDataset
    from multi_modal_csi_b200.train import _epoch_index_batches
    ds = TensorDataset(torch.arange(23).float().reshape(23, 1), torch.arange(23))
    torch.manual_seed(5)
    a = [b[1].tolist() for b in DataLoader(ds, 4, shuffle=True)] + [b[1].tolist() for b in DataLoader(ds, 4, shuffle=True)]
    torch.manual_seed(5)
    b = _epoch_index_batches(ds, 4, None) + _epoch_index_batches(ds, 4, None)
    assert a == b
    T, F = 12, 3
    lens = [12, 7, 1]
    chunks = [torch.rand(n, F) for n in lens]
    offs = [0, 12 * F, 19 * F]
    p = PackedCSIDataset(torch.cat([c.reshape(-1) for c in chunks]), offs, lens, F, torch.arange(3), T)
    for i, c in enumerate(chunks):
        xi, yi = p[i]
        assert xi.shape == (T, F) and torch.equal(xi[T - lens[i]:], c) and float(xi[:T - lens[i]].abs().sum()) == 0 and int(yi) == i
    with pytest.raises(ValueError):
        PackedCSIDataset(torch.zeros(13 * F), [0], [13], F, torch.zeros(1), T)


def test_cnn2d_module_contract_matches_reference_fixture(gold):
    """multi_modal_csi_b200.CNN_2D (the product module of the CSI-as-image path): the reference's 28 state_dict keys in
    its order, bit-identical initial weights under its seed, fp32 arena views, and no CPU path."""
    from multi_modal_csi_b200 import CNN_2D
    g = gold("cnn2d_anchor.npz")
    T, F, out, B = [int(v) for v in g["dims"]]
    torch.manual_seed(39)
    m = CNN_2D((T, F), (out,))
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    for k, v in sd.items():
        assert abs(v.double().sum().item() - float(g["init_sum/" + k])) <= 1e-6 * max(1.0, float(g["init_abs/" + k])), k
    assert m.geom.H == [300, 40, 9, 3] and m.geom.W == [270, 35, 7, 1] and m.geom.Kp == [736, 7200, 3136]
    assert sum(p.numel() for p in m.parameters()) == m.flat_params.numel() - sum(
        (-LY_numel(s)) % 4 for s in m.arena.shapes.values())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, T, F))
    with pytest.raises(ValueError):
        CNN_2D((20, 270), (54,))


def LY_numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def test_conv_and_head_bands_of_the_fused_gemms():
    """layout.StreamGeom.conv_bands / head_bands: the segments and output-column bands csi_gemm_nt_banded gets for the fused conv
    trio (that.py:122-135: kernel sizes 1/3/5 left, 1/2/3 right, padding="same") and the fused head convs (valid convolutions),
    and the stacked operand of build_pack_plan: every tap of every branch lands in the column block of its row shift, rows of a
    branch outside the taps it has stay zero."""
    import numpy as np
    from multi_modal_csi_b200 import layout as LY
    g = LY.ModelGeom(3000, 270, 54)
    for sg in g.streams:
        Dp = sg.Dp
        segs, bands = sg.conv_bands()
        shifts = [s[0] for s in segs]
        assert shifts == sorted({t - (k - 1) // 2 for k in sg.kernels for t in range(k)})
        for (sh, aoff, boff, klen), (lo, hi) in zip(segs, bands):
            assert aoff == 0 and klen == Dp and boff == shifts.index(sh) * Dp
            has = [j for j, k in enumerate(sg.kernels) if -((k - 1) // 2) <= sh <= k - 1 - (k - 1) // 2]
            assert (lo, hi) == (has[0] * Dp, (has[-1] + 1) * Dp) and lo % 16 == 0 and hi % 16 == 0
        hsegs, hbands = sg.head_bands()
        assert [s[0] for s in hsegs] == list(range(max(sg.head_k)))
        for t, (lo, hi) in enumerate(hbands):
            first = min(j for j, k in enumerate(sg.head_k) if k > t)
            assert (lo, hi) == (first * sg.head_np, len(sg.head_k) * sg.head_np)
    # the stacked forward operand against a direct construction
    arena = LY.build_arena(LY.parameter_specs(g))
    plan = LY.build_pack_plan(g, arena)
    from mirror_ops import MirrorOps
    ops = MirrorOps()
    params = torch.randn(arena.size)
    packed = torch.zeros(plan.size)
    ops.pack_weights(params, packed, ops.make_pack_table(plan.entries, "cpu"), len(plan.entries), plan.max_elems)
    sg, e = g.left, 2
    p, Dp, d = sg.prefix(e), sg.Dp, sg.d
    m = plan.mats["f:" + p + "layer_cnn"]
    W = packed[m.off:m.off + m.rows * m.ld].view(m.rows, m.ld)
    shifts = sg.conv_shifts
    want = torch.zeros_like(W)
    for j, k in enumerate(sg.kernels):
        name = f"{p}layer_cnn.{j}.0.weight"
        w = params[arena.offsets[name]:arena.offsets[name] + d * d * k].view(d, d, k)
        for t in range(k):
            c = shifts.index(t - (k - 1) // 2)
            want[j * Dp:j * Dp + d, c * Dp:c * Dp + d] = w[:, :, t]
    assert torch.equal(W, want)
