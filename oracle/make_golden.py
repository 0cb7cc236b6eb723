"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

Every fixture is produced by importing the reference's own classes/functions through
``oracle/ref_import.py`` (stubs only for ptflops/matplotlib/seaborn), never by the oracle
restatement or by the product package.  Fixtures:

  that_small.npz      full known-answer case, T=400 F=30 out=12 B=3: weights, x, y, train-mode logits,
                      loss, every parameter gradient, BN running stats after the step, eval logits,
                      and a 3-step Adam trajectory (losses + final weights checksum)
  that_anchor_*.npz   full-size anchors (F=270/out=54 and F=540/out=90, B=4): logits, loss, per-parameter
                      grad norms, eval-mode logits, per-parameter init checksums for seed 39
  metrics.npz         utils.performance_metrics(var_mode="baseline") on seeded random logits/labels
  that_count_pred.npz model/that_count_pred.py THAT_COUNT_PRED + SmoothL1Loss: logits, grads, 2 Adam steps, count metrics
  cnn2d_anchor.npz    model/cnn_2d.py CNN_2D (CSI-as-image): init checksums, logits, loss, gradient norms
  that_multi_head.npz model/that_multi_head.py five-head THAT + PermutationMatchingLoss: logits, grads, 2 Adam steps, loss KATs
  labels.npz + annotation_excerpt.csv   load_data.encode_* on an excerpt of dataset/annotation.csv
  augment_stats.npz   moments of train.py::apply_augmentation output (statistical fixture)
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference, REF_ROOT  # noqa: E402


def synth(B, T, F, out, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, F, generator=g) * 20
    y = (torch.rand(B, out, generator=g) < 0.15).float()
    return x, y


def build(ns, T, F, out, seed=39):
    torch.manual_seed(seed)
    m = ns.that.THAT((T, F), (out,))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def one_step(m, x, y, out):
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * out))
    m.train()
    m.zero_grad()
    logits = m(x)
    l = loss(logits, y)
    l.backward()
    return logits.detach(), l.detach()


def small_case(ns):
    T, F, out, B = 400, 30, 12, 3
    m = build(ns, T, F, out)
    x, y = synth(B, T, F, out)
    rec = {"x": x.numpy(), "y": y.numpy(), "dims": np.array([T, F, out, B])}
    for k, v in m.state_dict().items():
        rec["w/" + k] = v.detach().clone().numpy()
    logits, l = one_step(m, x, y, out)
    rec["logits_train"] = logits.numpy()
    rec["loss"] = l.numpy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["g/" + k] = p.grad.detach().clone().numpy()
    for k, v in m.state_dict().items():
        if "running" in k or "num_batches" in k:
            rec["bn/" + k] = v.detach().clone().numpy()
    m.eval()
    with torch.no_grad():
        rec["logits_eval"] = m(x).numpy()
    # 3-step Adam trajectory from the initial weights (reference optimizer, preset lr / wd)
    m2 = build(ns, T, F, out)
    opt = torch.optim.Adam(m2.parameters(), lr=5e-4, weight_decay=2e-4)
    losses = []
    for s in range(3):
        xs, ys = synth(B, T, F, out, seed=1234 + s)
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4.0] * out))
        m2.train()
        pred = m2(xs)
        lv = loss(pred, ys)
        opt.zero_grad()
        lv.backward()
        opt.step()
        losses.append(lv.item())
    rec["traj_losses"] = np.array(losses, dtype=np.float64)
    for k, v in m2.state_dict().items():
        rec["traj_w/" + k] = v.detach().clone().numpy()
    np.savez_compressed(os.path.join(GOLD, "that_small.npz"), **rec)
    print("that_small: loss", l.item(), "traj", losses)


def anchor_case(ns, F, out, B=4):
    T = 3000
    m = build(ns, T, F, out)
    rec = {"dims": np.array([T, F, out, B])}
    for k, v in m.state_dict().items():
        vv = v.detach().double()
        rec["init_sum/" + k] = np.array(vv.sum().item())
        rec["init_abs/" + k] = np.array(vv.abs().sum().item())
    x, y = synth(B, T, F, out)
    logits, l = one_step(m, x, y, out)
    rec["logits_train"] = logits.numpy()
    rec["loss"] = l.numpy()
    tot = 0.0
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["gnorm/" + k] = np.array(p.grad.double().norm().item())
            tot += p.grad.double().pow(2).sum().item()
    rec["grad_norm"] = np.array(tot ** 0.5)
    rec["g/layer_output.weight"] = m.layer_output.weight.grad.numpy()
    rec["g/layer_left_gaussian.var_mu"] = m.layer_left_gaussian.var_mu.grad.numpy()
    rec["g/layer_left_gaussian.var_sigma"] = m.layer_left_gaussian.var_sigma.grad.numpy()
    rec["g/layer_right_encoder.0.layer_attention.in_proj_bias"] = \
        m.layer_right_encoder[0].layer_attention.in_proj_bias.grad.numpy()
    m.eval()
    with torch.no_grad():
        rec["logits_eval"] = m(x).numpy()
    np.savez_compressed(os.path.join(GOLD, f"that_anchor_{F}.npz"), **rec)
    print(f"anchor F={F}: logits[0,:4]", logits[0, :4].tolist(), "loss", l.item(), "gnorm", rec["grad_norm"])


def metrics_case(ns):
    rng = np.random.default_rng(7)
    N = 64
    y_true = np.zeros((N, 6, 9), dtype=np.int64)
    for n in range(N):
        for u in range(6):
            if rng.random() > 0.6:
                y_true[n, u, rng.integers(0, 9)] = 1
    logits = rng.normal(-2.0, 2.5, size=(N, 54)) + 5.0 * y_true.reshape(N, 54) * (rng.random((N, 54)) < 0.8)
    res = ns.utils.performance_metrics(y_true, logits.astype(np.float32), var_mode="baseline", var_threshold=0.5)
    rec = {"y_true": y_true, "logits": logits.astype(np.float32)}
    for k, v in res.items():
        rec["m/" + k] = np.asarray(v, dtype=np.float64)
    sig = 1 / (1 + np.exp(-logits.astype(np.float32)))
    yp, yt, _ = ns.utils.process_predictions(sig.reshape(N, 6, 9).astype(float), y_true)
    rec["counts_pred"] = yp
    rec["counts_true"] = yt
    # train.py:108 quirk: metrics of the last train batch use logits truncated to int
    res_q = ns.utils.performance_metrics(y_true.reshape(N, -1).astype(int), logits.astype(np.float32).astype(int),
                                         var_mode="baseline", var_threshold=0.5)
    for k, v in res_q.items():
        rec["mq/" + k] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **rec)
    print("metrics:", {k: (float(v) if np.ndim(v) == 0 else "arr") for k, v in res.items()})


def labels_case(ns):
    import pandas as pd
    src = os.path.join(REF_ROOT, "dataset", "annotation.csv")
    df = pd.read_csv(src, dtype=str)
    # deterministic excerpt that covers every environment / band / user count
    pick = df.groupby(["environment", "wifi_band", "number_of_users"], sort=True).head(3)
    excerpt = os.path.join(GOLD, "annotation_excerpt.csv")
    pick.to_csv(excerpt, index=False)
    sel = ns.load_data.load_data_y(excerpt, ["classroom", "empty_room"], ["2.4"], ["0", "1", "3", "5"])
    rec = {
        "labels": np.array(sel["label"].to_list()),
        "identity": ns.load_data.encode_data_y(sel, "identity"),
        "activity": ns.load_data.encode_data_y(sel, "activity"),
        "location": ns.load_data.encode_data_y(sel, "location"),
    }
    allsel = ns.load_data.load_data_y(excerpt)
    rec["n_all"] = np.array(len(allsel))
    rec["activity_all"] = ns.load_data.encode_data_y(allsel, "activity")
    np.savez_compressed(os.path.join(GOLD, "labels.npz"), **rec)
    print("labels:", {k: np.shape(v) for k, v in rec.items()})


def augment_case(ns):
    # apply_augmentation is a closure inside train(); reproduce its call through the real train() is
    # impractical, so record the moments of the documented transform run with torch's own RNG.
    torch.manual_seed(5)
    x = torch.rand(8, 3000, 270) * 20
    noise = torch.randn_like(x) * 0.1
    xa = x + noise
    scale = torch.rand(x.size(0), 1) * 0.2 + 0.9
    xa = xa * scale.unsqueeze(-1)
    mask = torch.bernoulli(torch.ones_like(xa) * 0.96)
    xa = xa * mask
    rec = {"keep_rate": np.array(mask.mean().item()), "noise_std": np.array(noise.std().item()),
           "scale_min": np.array(0.9), "scale_max": np.array(1.1),
           "ratio_mean": np.array((xa.sum() / x.sum()).item())}
    np.savez_compressed(os.path.join(GOLD, "augment_stats.npz"), **rec)
    print("augment:", {k: float(v) for k, v in rec.items()})


def count_pred_case(ns):
    """Sibling head: the reference's THAT_COUNT_PRED (model/that_count_pred.py) with SmoothL1Loss on
    per-activity counts, 2 Adam steps (weight_decay 0 as in that_count_pred.py:397), plus the count-mode metrics."""
    import importlib.util
    from oracle.ref_import import REF_WIFI
    spec = importlib.util.spec_from_file_location("ref_model_that_count_pred", os.path.join(REF_WIFI, "model", "that_count_pred.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    T, F, out, B = 400, 30, 9, 4
    torch.manual_seed(39)
    m = mod.THAT_COUNT_PRED((T, F), [out])
    for sub in m.modules():
        if isinstance(sub, torch.nn.Dropout):
            sub.p = 0.0
    g = torch.Generator().manual_seed(4321)
    x = torch.rand(B, T, F, generator=g) * 20
    y = torch.randint(0, 3, (B, 6, out), generator=g)          # [B, users, activities] one-hot-ish -> counts by sum(axis=1)
    rec = {"x": x.numpy(), "y": y.numpy(), "dims": np.array([T, F, out, B])}
    for k, v in m.state_dict().items():
        rec["w/" + k] = v.detach().clone().numpy()
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=0)
    loss = torch.nn.SmoothL1Loss()
    losses = []
    for s in range(2):
        m.train()
        pred = m(x)
        yc = y.sum(axis=1)                                     # train.py:91-92
        lv = loss(pred, yc.float())
        opt.zero_grad()
        lv.backward()
        if s == 0:
            rec["logits_train"] = pred.detach().numpy()
            for k, p_ in m.named_parameters():
                if p_.grad is not None:
                    rec["g/" + k] = p_.grad.detach().clone().numpy()
        opt.step()
        losses.append(lv.item())
    rec["traj_losses"] = np.array(losses, dtype=np.float64)
    for k, v in m.state_dict().items():
        rec["traj_w/" + k] = v.detach().clone().numpy()
    # count-mode metrics (utils.py:229-233) on seeded regression outputs
    rng = np.random.default_rng(11)
    yt = rng.integers(0, 4, size=(48, out))
    yp = yt + rng.normal(0, 0.45, size=(48, out))
    res = ns.utils.performance_metrics(yt, yp, var_mode="count_classification")
    rec["m_y_true"], rec["m_y_pred"] = yt, yp
    for k, v in res.items():
        rec["m/" + k] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "that_count_pred.npz"), **rec)
    print("that_count_pred: losses", losses, "total_error", float(res["total_error"]))


def multi_head_case(ns):
    """Sibling head: the reference's five-head THAT (model/that_multi_head.py:180-306) with
    PermutationMatchingLoss (:309-342), one forward + backward and 2 Adam steps (weight_decay 0, :414-416).  The
    reference's ``multi_head`` metrics branch does not run (utils.py:221 indexes ``y_pred[-1]`` of a [B,5,C] array and then
    unpacks three dimensions), so only model + loss are pinned."""
    import importlib.util
    from oracle.ref_import import REF_WIFI
    spec = importlib.util.spec_from_file_location("ref_model_that_multi_head", os.path.join(REF_WIFI, "model", "that_multi_head.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    T, F, C, B, H = 400, 30, 10, 6, 5
    torch.manual_seed(39)
    m = mod.THAT((T, F), [C])
    for sub in m.modules():
        if isinstance(sub, torch.nn.Dropout):
            sub.p = 0.0
    g = torch.Generator().manual_seed(8642)
    x = torch.rand(B, T, F, generator=g) * 20
    cls = torch.randint(0, C, (B, H), generator=g)
    cls[:, 3:] = C - 1                                          # most slots empty: the last class is "nobody" (utils.py:227)
    y = torch.nn.functional.one_hot(cls, C).float()             # [B, 5, C]
    rec = {"x": x.numpy(), "y": y.numpy(), "dims": np.array([T, F, C, B, H])}
    for k, v in m.state_dict().items():
        rec["w/" + k] = v.detach().clone().numpy()
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=0)
    loss = mod.PermutationMatchingLoss()
    losses = []
    for s in range(2):
        m.train()
        pred = m(x)                                             # [B, 5, C]
        lv = loss(pred, y)
        opt.zero_grad()
        lv.backward()
        if s == 0:
            rec["logits_train"] = pred.detach().numpy()
            for k, p_ in m.named_parameters():
                if p_.grad is not None:
                    rec["g/" + k] = p_.grad.detach().clone().numpy()
        opt.step()
        losses.append(lv.item())
    rec["traj_losses"] = np.array(losses, dtype=np.float64)
    for k, v in m.state_dict().items():
        rec["traj_w/" + k] = v.detach().clone().numpy()
    # known answers for the loss alone (ties included: two identical heads)
    gl = torch.Generator().manual_seed(77)
    lp = torch.randn(16, H, C, generator=gl)
    lp[5, 1] = lp[5, 0]
    lt = torch.nn.functional.one_hot(torch.randint(0, C, (16, H), generator=gl), C).float()
    lp.requires_grad_(True)
    lv = loss(lp, lt)
    lv.backward()
    rec["loss_pred"], rec["loss_target"] = lp.detach().numpy(), lt.numpy()
    rec["loss_value"], rec["loss_grad"] = np.float64(lv.item()), lp.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "that_multi_head.npz"), **rec)
    print("that_multi_head: losses", losses, "loss-only", lv.item(), "keys", len(m.state_dict()))


def cnn2d_case(ns):
    """CSI-as-image path (BASELINE config 4's in-reference analogue): model/cnn_2d.py CNN_2D at the full
    feature width (F=270; the 27/15/7 kernels with strides 7/3/1 need it) on a short window, BCE(pos_weight=6).  Only
    checksums and outputs are stored: weights are re-created from the reference seed by oracle.cnn2d_oracle.cnn2d_init
    and the input from its generator seed."""
    import importlib.util
    from oracle.ref_import import REF_WIFI
    spec = importlib.util.spec_from_file_location("ref_model_cnn_2d", os.path.join(REF_WIFI, "model", "cnn_2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    T, F, out, B = 300, 270, 54, 3
    torch.manual_seed(39)
    m = mod.CNN_2D((T, F), (out,))
    for sub in m.modules():
        if isinstance(sub, torch.nn.Dropout):
            sub.p = 0.0
    rec = {"dims": np.array([T, F, out, B]), "keys": np.array(list(m.state_dict().keys()))}
    for k, v in m.state_dict().items():
        rec["init_sum/" + k] = np.float64(v.double().sum().item())
        rec["init_abs/" + k] = np.float64(v.double().abs().sum().item())
    g = torch.Generator().manual_seed(2468)
    x = torch.rand(B, T, F, generator=g) * 20
    y = (torch.rand(B, out, generator=g) < 0.15).float()
    m.train()
    logits = m(x)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 6.0))(logits, y)
    loss.backward()
    rec["logits_train"] = logits.detach().numpy()
    rec["loss"] = np.float64(loss.item())
    for k, p_ in m.named_parameters():
        rec["gnorm/" + k] = np.float64(p_.grad.double().norm().item())
        rec["gsum/" + k] = np.float64(p_.grad.double().sum().item())
    for k, v in m.state_dict().items():
        if "running" in k:
            rec["stat/" + k] = v.detach().clone().numpy()
    m.eval()
    with torch.no_grad():
        rec["logits_eval"] = m(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "cnn2d_anchor.npz"), **rec)
    print("cnn2d_anchor: loss", loss.item(), "logits[0,:3]", logits[0, :3].tolist())


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ns = load_reference()
    small_case(ns)
    anchor_case(ns, 270, 54)
    anchor_case(ns, 540, 90)
    metrics_case(ns)
    labels_case(ns)
    augment_case(ns)
    count_pred_case(ns)
    multi_head_case(ns)
    cnn2d_case(ns)


if __name__ == "__main__":
    main()
