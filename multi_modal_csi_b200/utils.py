"""Metrics, result encoding and checkpoint helpers with the semantics of benchmark/wifi_csi/utils.py.

These run on the CPU in numpy once per epoch, exactly where the reference runs them (train.py:105-127); they are
not on the accelerated path.  ``var_mode="baseline"`` (BCE / THAT) and ``"count_classification"`` (SmoothL1 /
THAT_COUNT_PRED) are implemented.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


class NumpyEncoder(json.JSONEncoder):
    """utils.py:185-193."""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def process_predictions(y_pred, y_true, var_threshold=0.5):
    """utils.py:147-183: per user, the arg-max class counts iff its probability exceeds the threshold; classes are
    then summed over users.  y_pred, y_true: [N, users, classes] -> ([N, classes], [N, classes], N)."""
    y_pred = np.asarray(y_pred)
    top = y_pred.argmax(axis=2)
    top_p = np.take_along_axis(y_pred, top[..., None], axis=2)[..., 0]
    hit = np.zeros_like(y_pred)
    np.put_along_axis(hit, top[..., None], (top_p > var_threshold)[..., None].astype(y_pred.dtype), axis=2)
    return hit.sum(axis=1), np.asarray(y_true).sum(axis=1), y_true.shape[0]


def error_per_number_person(y_pred, y_true):
    """utils.py:103-120: mean absolute count error over the samples with 1..5 people present."""
    people = y_true.sum(axis=1)
    err = np.abs(y_pred - y_true).sum(axis=1)
    with np.errstate(invalid="ignore"), warnings_off():
        return [err[people == n].mean() if np.any(people == n) else float("nan") for n in range(1, 6)]


class warnings_off:
    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self._cm.__exit__(*a)


def count_error(y_pred, y_true):
    """utils.py:122-135: |#people predicted - #people present| per sample."""
    return np.abs(y_pred.sum(axis=1) - y_true.sum(axis=1))


def calculate_scores(y_true, y_pred):
    """utils.py:196-211: count-based per-activity precision / recall / F1 / accuracy, macro-averaged."""
    tp = np.minimum(y_true, y_pred).sum(axis=0)
    tn = (np.maximum(y_true, y_pred) == 0).astype(np.int64).sum(axis=0)
    fp = np.maximum(0, y_pred - y_true).sum(axis=0)
    fn = np.maximum(0, y_true - y_pred).sum(axis=0)
    precision = np.where((tp + fp) > 0, tp / (tp + fp + 1e-6), 0)
    recall = np.where((tp + fn) > 0, tp / (tp + fn + 1e-6), 0)
    f1 = np.where((precision + recall) > 0, 2 * (precision * recall) / (precision + recall + 1e-6), 0)
    acc = (tp + tn) / (tp + fn + tn + fp)
    return precision.mean(), recall.mean(), f1.mean(), acc.mean()


def threshold_round(x, threshold=0.3):
    """utils.py:137-145 (vectorised): round up iff the fractional part exceeds ``threshold``."""
    x = np.asarray(x, dtype=float)
    fl = np.floor(x)
    return np.where(x - fl > threshold, np.ceil(x), fl)


def performance_metrics(y_true, y_pred, var_mode="baseline", var_threshold=0.5):
    """utils.py:213-270 for var_mode "baseline", "count_classification" and "multi_head".  In baseline mode the decision threshold
    is the reference's hard-coded 0.5 (utils.py:238 ignores ``var_threshold``) and classes per user are fixed at 9
    (utils.py:236); in count mode predictions are threshold-rounded at 0.5 and clipped to [0, 5] (utils.py:229-233)."""
    y_true = np.array(y_true)
    y_pred = np.array(y_pred)
    if var_mode == "multi_head":
        # utils.py:220-228: per head the arg-max class, one-hot, summed over the heads = per-class counts; the last class
        # ("nobody") is dropped.  The reference first takes ``y_pred[-1]`` (it expects a stack of per-stage outputs
        # [S, B, heads, classes]; with the [B, heads, classes] array its own model returns, that line breaks the unpacking
        # that follows): a 4-D input is reduced the reference's way, a 3-D one is used as it is.
        if y_pred.ndim == 4:
            y_pred = y_pred[-1]
        if y_pred.ndim != 3 or y_true.ndim != 3:
            raise ValueError(f"multi_head metrics need [B, heads, classes] predictions and targets, got {y_pred.shape} / {y_true.shape}")
        num_classes = y_pred.shape[-1]
        y_pred = np.eye(num_classes)[np.argmax(y_pred, axis=-1)].sum(axis=1)[:, :-1]
        y_true = y_true.sum(axis=1)[:, :-1]
    elif var_mode == "count_classification":
        y_pred = np.clip(threshold_round(y_pred, threshold=0.5), 0, 5)
    elif var_mode == "baseline":
        y_pred = (1 / (1 + np.exp(-y_pred))).astype(float)
        y_true = y_true.reshape(y_true.shape[0], -1, 9)
        y_pred = y_pred.reshape(y_true.shape)
        y_pred, y_true, _ = process_predictions(y_pred, y_true, var_threshold=0.5)
    else:
        raise ValueError(f"Unsupported var_mode: {var_mode}")
    n = y_true.shape[0]
    diff = np.abs(y_true - y_pred)
    counting = count_error(y_pred, y_true)
    precision, recall, f1, acc = calculate_scores(y_true, y_pred)
    return {
        "total_error": diff.sum() / n,
        "perfect_prediction_percentage": (np.all(diff == 0, axis=1).sum() / n) * 100,
        "accuracy": acc,
        "error_per_person": error_per_number_person(y_pred, y_true),
        "mean_count_error": counting.mean(),
        "counting_error_perPerson": counting,
        "precision": precision,
        "recall": recall,
        "f1_score": f1,
    }


def reduce_dataset(data, num_object_queries=None):
    """utils.py:272-287: labels [N, 6 users, 9 activities] -> [N, 5 slots, 10 classes] for the multi-head sibling: the first
    all-zero user row is dropped, a "nobody" class column is appended and set on the remaining empty rows (optionally
    padded with "nobody" rows up to ``num_object_queries``)."""
    nobody = np.zeros(data.shape[-1] + 1)
    nobody[-1] = 1
    out = []
    for sample in data:
        new = np.delete(sample, (sample.sum(axis=1) == 0).argmax(), axis=0)
        new = np.hstack((new, np.zeros((new.shape[0], 1))))
        new[new.sum(axis=1) == 0, :] = nobody
        if num_object_queries:
            new = np.concatenate((new, np.repeat([nobody], num_object_queries - new.shape[0], axis=0)))
        out.append(new)
    return np.array(out)


def save_model_components(preset, model):
    """utils.py:89-101: ``{saving_path}model_0/PT_{envs}_{model}.pth`` holding ``model.state_dict()``."""
    save_dir = preset.get("saving_path") + "model_0"
    os.makedirs(save_dir, exist_ok=True)
    env = "_".join(preset["data"]["environment"])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}     # plain tensors, no arena views
    torch.save(sd, f"{save_dir}/PT_{env}_{preset['model']}.pth")


def load_model_components(model, load_path, lr, scenario="full", device=None):
    """utils.py:16-86.  THAT has no feature_extractor/encoder/decoder sub-modules, so only "full" applies."""
    if scenario != "full":
        raise ValueError(f"transfer scenario {scenario!r} does not apply to THAT (no such sub-modules)")
    state = torch.load(load_path, map_location="cpu")
    model.load_state_dict(state)
    return model, [{"params": list(model.parameters()), "lr": lr}]
