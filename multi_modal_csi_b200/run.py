"""``run_that`` (benchmark/wifi_csi/model/that.py:307-499) and the CLI driver of run_main.py:20-168 for model THAT."""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np
import torch
from sklearn.model_selection import train_test_split
from torch.utils.data import TensorDataset

from .load_data import encode_data_y, load_data_x, load_data_y
from .optim import FusedAdam
from .preset import preset
from .that import THAT, THAT_COUNT_PRED, THAT_MULTI_HEAD, PermutationMatchingLoss
from .train import _log, _wandb, train
from .utils import NumpyEncoder, load_model_components, performance_metrics, reduce_dataset, save_model_components


def run_that(data_train_x, data_train_y, data_test_x, data_test_y, var_repeat=10):
    """Train and evaluate THAT ``var_repeat`` times; returns the metrics dict of the last repeat (that.py:499)."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    data_valid_x, data_test_x, data_valid_y, data_test_y = train_test_split(
        data_test_x, data_test_y, test_size=0.5, shuffle=True, random_state=39)                     # that.py:332-335
    data_valid_x = data_valid_x.reshape(data_valid_x.shape[0], data_valid_x.shape[1], -1)
    data_train_x = data_train_x.reshape(data_train_x.shape[0], data_train_x.shape[1], -1)
    data_test_x = data_test_x.reshape(data_test_x.shape[0], data_test_x.shape[1], -1)
    var_x_shape, var_y_shape = data_train_x[0].shape, data_train_y[0].reshape(-1).shape
    data_train_set = TensorDataset(torch.from_numpy(data_train_x), torch.from_numpy(data_train_y))
    data_valid_set = TensorDataset(torch.from_numpy(data_valid_x), torch.from_numpy(data_valid_y))
    n_params = sum(p.numel() for p in THAT(var_x_shape, var_y_shape).parameters())
    print("Parameters:", n_params)
    results = {k: [] for k in ("accuracy", "time_train", "time_test", "total_error", "precision", "recall", "f1_score")}
    dict_true_acc = None
    for var_r in range(var_repeat):
        print("Repeat", var_r)
        if _wandb is not None:
            _wandb.init(project="final_results", name=f"BCE_THAT_{var_r}_" + "_".join(preset["data"]["environment"]),
                        config=preset, reinit=True)
        torch.random.manual_seed(var_r + 39)                                                         # that.py:381
        model_that = THAT(var_x_shape, var_y_shape, act_dtype=preset["nn"].get("dtype", "bf16"),
                          max_batch=preset["nn"]["batch_size"]).to(device)
        if preset.get("pretrained_path"):
            model_that, param_groups = load_model_components(model_that, preset["pretrained_path"], preset["nn"]["lr"],
                                                             preset.get("transfer_scenario"), device)
            optimizer = FusedAdam(param_groups[0]["params"], lr=param_groups[0]["lr"])
        else:
            optimizer = FusedAdam(model_that.parameters(), lr=preset["nn"]["lr"], weight_decay=preset["nn"]["weight_decay"])
        loss_mode = "baseline"
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([4] * var_y_shape[-1]).to(device))  # that.py:401
        var_time_0 = time.time()
        var_best_weight = train(model=model_that, optimizer=optimizer, loss=loss, data_train_set=data_train_set,
                                data_test_set=data_valid_set, var_threshold=preset["nn"]["threshold"],
                                var_batch_size=preset["nn"]["batch_size"], var_epochs=preset["nn"]["epoch"], device=device,
                                var_mode=loss_mode)
        var_time_1 = time.time()
        if preset.get("save_model"):
            save_model_components(preset, model_that)
        model_that.load_state_dict(var_best_weight)
        with torch.no_grad():                         # the model is still in eval() from train(), as in that.py:427-432
            predict_test_y = model_that(torch.from_numpy(data_test_x).to(device))
        predict_test_y = predict_test_y.detach().cpu().numpy()
        var_time_2 = time.time()
        dict_true_acc = performance_metrics(data_test_y, predict_test_y, var_mode=loss_mode,
                                            var_threshold=preset["nn"]["threshold"])
        payload = {"repeat": var_r, "train_time": var_time_1 - var_time_0, "test_time": var_time_2 - var_time_1,
                   "TOTAL_TESTSET_ERROR": dict_true_acc["total_error"],
                   "TOTAL_TESTSET_perfect_prediction_percentage": dict_true_acc["perfect_prediction_percentage"],
                   "TOTAL_ACCURACY": dict_true_acc["accuracy"], "mean_count_error": dict_true_acc["mean_count_error"],
                   "precision": dict_true_acc["precision"], "recall": dict_true_acc["recall"],
                   "f1_score": dict_true_acc["f1_score"]}
        for i in range(5):
            payload[f"error_per_person_{i + 1}"] = dict_true_acc["error_per_person"][i]
        _log(payload)
        print(" %.6fs" % (time.time() - var_time_1), "- Total Error %.6f" % dict_true_acc["total_error"],
              "-  perfect_prediction_percentage %.6f" % dict_true_acc["perfect_prediction_percentage"])
        results["accuracy"].append(dict_true_acc["perfect_prediction_percentage"])
        results["time_train"].append(var_time_1 - var_time_0)
        results["time_test"].append(var_time_2 - var_time_1)
        for k in ("total_error", "precision", "recall", "f1_score"):
            results[k].append(dict_true_acc[k])
    _log({"avg_" + ("train_time" if k == "time_train" else "test_time" if k == "time_test" else k): sum(v) / len(v)
          for k, v in results.items() if v})
    if _wandb is not None and getattr(_wandb, "run", None) is not None:
        _wandb.finish()
    return dict_true_acc


def master_splitter(preset, var_task, var_model, var_users):
    """run_main.py:20-66: per-environment 80/20 split (random_state=103) so no environment leaks across the split."""
    xs_tr, xs_te, ys_tr, ys_te = [], [], [], []
    for env in preset["data"]["environment"]:
        data_pd_y = load_data_y(preset["path"]["data_y"], var_environment=[env], var_wifi_band=preset["data"]["wifi_band"],
                                var_num_users=var_users)
        X = load_data_x(preset["path"]["data_x"], data_pd_y["label"].to_list())
        y = encode_data_y(data_pd_y, var_task)
        if var_model == "THAT_MULTI_HEAD":
            y = reduce_dataset(y)                                          # run_main.py:39-40
        X_train, X_test, y_train, y_test = train_test_split(X, y, test_size=0.2, shuffle=True, random_state=103)
        xs_tr.append(X_train); xs_te.append(X_test); ys_tr.append(y_train); ys_te.append(y_test)
    return (np.concatenate(xs_tr, 0), np.concatenate(xs_te, 0), np.concatenate(ys_tr, 0), np.concatenate(ys_te, 0))


def parse_args():
    a = argparse.ArgumentParser()
    a.add_argument("--model", default=preset["model"], type=str)
    a.add_argument("--task", default=preset["task"], type=str)
    a.add_argument("--repeat", default=preset["repeat"], type=int)
    a.add_argument("--users", default="0, 1,2,3,4,5", type=str, help="Comma-separated list of user IDs")
    return a.parse_args()


def run_that_multihead(data_train_x, data_train_y, data_test_x, data_test_y, var_repeat=10):
    """model/that_multi_head.py:345-482: the five-head THAT trained with PermutationMatchingLoss (Adam without weight decay,
    per-step cosine schedule with warm-up, ``var_mode="multi_head"``); labels are ``[N, 5, classes]`` one-hot slots
    (``utils.reduce_dataset``).  Returns the metrics dict of the last repeat.  The reference evaluates with a helper that
    does not exist in its utils.py (``calculate_matrix_absolute_error``, that_multi_head.py:461); the count metrics of
    ``performance_metrics(var_mode="multi_head")`` (utils.py:220-228) are what that call is meant to produce."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    data_train_x = data_train_x.reshape(data_train_x.shape[0], data_train_x.shape[1], -1)
    data_test_x = data_test_x.reshape(data_test_x.shape[0], data_test_x.shape[1], -1)
    var_x_shape, var_y_shape = data_train_x[0].shape, [data_train_y[0].shape[1]]                  # that_multi_head.py:374
    data_train_set = TensorDataset(torch.from_numpy(data_train_x), torch.from_numpy(data_train_y))
    data_test_set = TensorDataset(torch.from_numpy(data_test_x), torch.from_numpy(data_test_y))
    dict_true_acc, result_accuracy = None, []
    for var_r in range(var_repeat):
        print("Repeat", var_r)
        if _wandb is not None:
            _wandb.init(project="wifi-based-model-THAT_ENCODER", name=f"Repeat_{var_r}",
                        config={"model": "THAT_MultiHead", "repeat": var_r}, reinit=True)
        torch.random.manual_seed(var_r + 39)
        model_that = THAT_MULTI_HEAD(var_x_shape, var_y_shape, act_dtype=preset["nn"].get("dtype", "bf16"),
                                     max_batch=preset["nn"]["batch_size"]).to(device)
        optimizer = FusedAdam(model_that.parameters(), lr=preset["nn"]["lr"], weight_decay=0)
        loss = PermutationMatchingLoss()
        var_mode = "multi_head"
        var_time_0 = time.time()
        var_best_weight = train(model=model_that, optimizer=optimizer, loss=loss, data_train_set=data_train_set,
                                data_test_set=data_test_set, var_threshold=preset["nn"]["threshold"],
                                var_batch_size=preset["nn"]["batch_size"], var_epochs=preset["nn"]["epoch"], device=device,
                                var_mode=var_mode)
        var_time_1 = time.time()
        model_that.load_state_dict(var_best_weight)
        with torch.no_grad():
            predict_test_y = model_that(torch.from_numpy(data_test_x).to(device))
        predict_test_y = predict_test_y.detach().cpu().numpy()
        var_time_2 = time.time()
        dict_true_acc = performance_metrics(data_test_y, predict_test_y, var_mode=var_mode)
        _log({"repeat": var_r, "train_time": var_time_1 - var_time_0, "test_time": var_time_2 - var_time_1,
              "TOTAL_TESTSET_ERROR": dict_true_acc["total_error"],
              "TOTAL_TESTSET_perfect_prediction_percentage": dict_true_acc["perfect_prediction_percentage"],
              "TOTAL_ACCURACY": dict_true_acc["accuracy"]})
        print(" %.6fs" % (time.time() - var_time_1), "- Total Error %.6f" % dict_true_acc["total_error"],
              "-  perfect_prediction_percentage %.6f" % dict_true_acc["perfect_prediction_percentage"])
        result_accuracy.append(dict_true_acc["perfect_prediction_percentage"])
    if result_accuracy:
        _log({"avg_accuracy": sum(result_accuracy) / len(result_accuracy)})
    if _wandb is not None and getattr(_wandb, "run", None) is not None:
        _wandb.finish()
    return dict_true_acc


def run_that_count_pred(data_train_x, data_train_y, data_test_x, data_test_y, var_repeat=10):
    """model/that_count_pred.py:334-509: THAT_COUNT_PRED trained with SmoothL1 on per-activity counts (Adam without
    weight decay, ``var_mode="count_classification"``); returns the metrics dict of the last repeat."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    data_valid_x, data_test_x, data_valid_y, data_test_y = train_test_split(
        data_test_x, data_test_y, test_size=0.5, shuffle=True, random_state=39)
    data_valid_x = data_valid_x.reshape(data_valid_x.shape[0], data_valid_x.shape[1], -1)
    data_train_x = data_train_x.reshape(data_train_x.shape[0], data_train_x.shape[1], -1)
    data_test_x = data_test_x.reshape(data_test_x.shape[0], data_test_x.shape[1], -1)
    var_x_shape, var_y_shape = data_train_x[0].shape, [data_train_y[0].shape[1]]                  # that_count_pred.py:366
    data_train_set = TensorDataset(torch.from_numpy(data_train_x), torch.from_numpy(data_train_y))
    data_valid_set = TensorDataset(torch.from_numpy(data_valid_x), torch.from_numpy(data_valid_y))
    dict_true_acc = None
    for var_r in range(var_repeat):
        print("Repeat", var_r)
        if _wandb is not None:
            _wandb.init(project="final_results", name=f"DEM_THAT_{var_r}_" + "_".join(preset["data"]["environment"]),
                        config=preset, reinit=True)
        torch.random.manual_seed(var_r + 39)
        model_that = THAT_COUNT_PRED(var_x_shape, var_y_shape, act_dtype=preset["nn"].get("dtype", "bf16"),
                                     max_batch=preset["nn"]["batch_size"]).to(device)
        optimizer = FusedAdam(model_that.parameters(), lr=preset["nn"]["lr"], weight_decay=0)
        loss_mode = "count_classification"
        loss = torch.nn.SmoothL1Loss()
        var_time_0 = time.time()
        var_best_weight = train(model=model_that, optimizer=optimizer, loss=loss, data_train_set=data_train_set,
                                data_test_set=data_valid_set, var_threshold=preset["nn"]["threshold"],
                                var_batch_size=preset["nn"]["batch_size"], var_epochs=preset["nn"]["epoch"], device=device,
                                var_mode=loss_mode)
        var_time_1 = time.time()
        model_that.load_state_dict(var_best_weight)
        with torch.no_grad():
            predict_test_y = model_that(torch.from_numpy(data_test_x).to(device))
        predict_test_y = predict_test_y.detach().cpu().numpy()
        var_time_2 = time.time()
        dict_true_acc = performance_metrics(data_test_y.sum(axis=1), predict_test_y, var_mode=loss_mode)
        payload = {"repeat": var_r, "train_time": var_time_1 - var_time_0, "test_time": var_time_2 - var_time_1,
                   "TOTAL_TESTSET_ERROR": dict_true_acc["total_error"],
                   "TOTAL_TESTSET_perfect_prediction_percentage": dict_true_acc["perfect_prediction_percentage"],
                   "TOTAL_ACCURACY": dict_true_acc["accuracy"], "mean_count_error": dict_true_acc["mean_count_error"],
                   "precision": dict_true_acc["precision"], "recall": dict_true_acc["recall"],
                   "f1_score": dict_true_acc["f1_score"]}
        for i in range(5):
            payload[f"error_per_person_{i + 1}"] = dict_true_acc["error_per_person"][i]
        _log(payload)
        print(" %.6fs" % (time.time() - var_time_1), "- Total Error %.6f" % dict_true_acc["total_error"],
              "-  perfect_prediction_percentage %.6f" % dict_true_acc["perfect_prediction_percentage"])
    if _wandb is not None and getattr(_wandb, "run", None) is not None:
        _wandb.finish()
    return dict_true_acc


def run():
    """run_main.py:88-160 restricted to the models on this path: THAT and its siblings THAT_COUNT_PRED / THAT_MULTI_HEAD."""
    var_args = parse_args()
    var_users = [u.strip() for u in var_args.users.split(",")]
    preset["repeat"] = 1 if not preset["pretrained_path"] else preset["repeat"]
    if var_args.model not in ("THAT", "THAT_COUNT_PRED", "THAT_MULTI_HEAD"):
        raise Exception("Not valid name for model")                       # run_main.py:140
    data_train_x, data_test_x, data_train_y, data_test_y = master_splitter(preset, var_args.task, var_args.model, var_users)
    runner = {"THAT": run_that, "THAT_COUNT_PRED": run_that_count_pred, "THAT_MULTI_HEAD": run_that_multihead}[var_args.model]
    result = runner(data_train_x, data_train_y, data_test_x, data_test_y, var_args.repeat)
    result["model"], result["task"], result["data"], result["nn"] = var_args.model, var_args.task, preset["data"], preset["nn"]
    print(result)
    os.makedirs(os.path.dirname(preset["path"]["save"]) or ".", exist_ok=True)
    with open(preset["path"]["save"], "w") as var_file:
        json.dump(result, var_file, indent=4, cls=NumpyEncoder)


if __name__ == "__main__":
    run()
