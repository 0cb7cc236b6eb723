"""``train(...)`` with the signature and semantics of benchmark/wifi_csi/train.py:36-176.

Kept from the reference: shuffled DataLoader with pinned memory (:48), the LAST batch of every epoch is skipped
(:81-82), augmentation only while ``model.training`` (:88-89), ``baseline`` labels flattened to [B, -1] (:93-94),
``count_classification`` labels summed over users to per-activity counts (:91-92, :116-117), the
once-per-epoch metrics of the last train batch on int-truncated logits (:105-109), whole-validation-set evaluation
in one batch (:49,111-127), the wandb keys (:130-144), the selection rule f1 AND perfect-prediction-percentage
(:159-166) and early stopping after ``patience`` non-improving epochs (:167-174).

Changed underneath: when the model is a multi_modal_csi_b200.THAT on a CUDA device, the optimizer a FusedAdam and the
loss a uniform-pos_weight BCEWithLogitsLoss, one step is ``model.fused_train_step`` (augmentation, forward, loss,
backward and Adam in hand-written sm_100a kernels) and the batches come from ``loader.CSIBatchSource``: the dataset
(dense TensorDataset or the ragged ``PackedCSIDataset``) is either resident in HBM (a batch = offsets into it) or
streamed from page-locked host memory, double-buffered on a copy stream, instead of the reference's per-batch host
collate + pin + ``.to(device)`` (train.py:48,84-86).  Otherwise the same loop runs through autograd on the model's
kernels with torch's loss/optimizer.  Under ``torch.distributed`` the batches are sharded over ranks, every rank starts
from rank 0's weights, gradients are all-reduced before the update and BatchNorm running statistics are taken from rank 0
before every evaluation, so that the selection / early-stopping decisions are identical on all ranks.
"""
from __future__ import annotations

import time
from copy import deepcopy

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, TensorDataset
from torch.utils.data.distributed import DistributedSampler

from .loader import CSIBatchSource, PackedCSIDataset
from .optim import FusedAdam
from .parallel import GradSync, bind_to_gpu_numa_node, broadcast_buffers, broadcast_parameters
from .arena_module import ArenaModule
from .cnn2d import CNN_2D
from .that import PermutationMatchingLoss
from .that import THAT
from .utils import performance_metrics

try:                                    # wandb is optional; WANDB_MODE=disabled also works as in the reference
    import wandb as _wandb
except Exception:                       # pragma: no cover
    _wandb = None


def _log(payload):
    if _wandb is not None and getattr(_wandb, "run", None) is not None:
        _wandb.log(payload)


def apply_augmentation(x_batch: torch.Tensor) -> torch.Tensor:
    """train.py:65-73 with torch ops (used on the generic path; the fused path does this inside csi_pool_dual)."""
    x_batch = x_batch + torch.randn_like(x_batch) * 0.1
    scale = torch.rand(x_batch.size(0), 1, device=x_batch.device) * 0.2 + 0.9
    x_batch = x_batch * scale.unsqueeze(-1)
    return x_batch * torch.bernoulli(torch.ones_like(x_batch) * 0.96)


def _uniform_pos_weight(loss):
    if not isinstance(loss, torch.nn.BCEWithLogitsLoss) or loss.reduction != "mean" or loss.weight is not None:
        return None
    pw = loss.pos_weight
    if pw is None:
        return 1.0
    v = float(pw.flatten()[0])
    return v if bool((pw == v).all()) else None


def cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, min_lr_ratio=0.1):
    """train.py:26-33 (``get_cosine_schedule_with_warmup``): LambdaLR, linear warm-up then cosine decay floored at
    ``min_lr_ratio``.  FusedAdam reads ``param_groups[0]["lr"]`` on the host at every step, so the schedule needs no kernel."""
    import math

    def lr_lambda(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        progress = float(current_step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(min_lr_ratio, 0.5 * (1.0 + math.cos(math.pi * progress)))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda)


def _epoch_index_batches(dataset, batch_size, sampler):
    """The index lists a ``DataLoader(dataset, batch_size, shuffle=True)`` would collate this epoch (same RNG draws:
    the loader iterator's base seed first, then the RandomSampler's seed), or the rank's shard under a DistributedSampler."""
    if sampler is not None:
        idx = list(iter(sampler))
    else:
        torch.empty((), dtype=torch.int64).random_()                     # _BaseDataLoaderIter draws its base seed first
        idx = list(iter(torch.utils.data.RandomSampler(dataset)))
    return [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]


def train(model, optimizer, loss, data_train_set: TensorDataset, data_test_set: TensorDataset, var_threshold: float,
          var_batch_size: int, var_epochs: int, device, var_mode: str, patience: int = 150, loader_mode: str = "auto"):
    device = torch.device(device)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    sampler = DistributedSampler(data_train_set, shuffle=True, drop_last=False) if distributed else None
    data_test_loader = DataLoader(data_test_set, len(data_test_set))
    pos_weight = _uniform_pos_weight(loss)
    loss_kind = None                                    # which fused loss kernel implements ``loss``
    if pos_weight is not None and var_mode == "baseline":
        loss_kind = "bce"
    elif (isinstance(loss, torch.nn.SmoothL1Loss) and loss.reduction == "mean" and float(loss.beta) == 1.0
          and var_mode == "count_classification"):
        loss_kind, pos_weight = "smooth_l1", 1.0
    elif isinstance(loss, PermutationMatchingLoss) and var_mode == "multi_head":
        loss_kind, pos_weight = "perm_ce", 1.0
    fused = (isinstance(model, (THAT, CNN_2D)) and isinstance(optimizer, FusedAdam) and loss_kind is not None
             and device.type == "cuda" and (loss_kind == "bce" or isinstance(model, THAT)))
    sync = GradSync(model, dist.get_world_size()) if distributed and isinstance(model, ArenaModule) else None
    if distributed and isinstance(model, ArenaModule):
        if device.type == "cuda":
            bind_to_gpu_numa_node(device)               # before the loader page-locks the dataset: node-local host memory
        broadcast_parameters(model)                     # every rank starts from rank 0's weights and BatchNorm buffers
        model.rng_seed = (model.rng_seed + 7919 * dist.get_rank()) & 0x7FFFFFFF      # decorrelated dropout / augmentation
    source = data_train_loader = None
    if fused:
        source = CSIBatchSource(data_train_set, device, var_batch_size, mode=loader_mode)
        n_train = len(sampler) if sampler is not None else len(data_train_set)
        total_batches = (n_train + var_batch_size - 1) // var_batch_size
    else:
        if isinstance(data_train_set, PackedCSIDataset):
            raise ValueError("PackedCSIDataset feeds the fused path (THAT + FusedAdam + a fused loss on a CUDA device)")
        data_train_loader = DataLoader(data_train_set, var_batch_size, shuffle=sampler is None, sampler=sampler, pin_memory=True)
        total_batches = len(data_train_loader)

    var_best_f1_score, var_best_PPP, var_best_weight, counter = 0, 0, None, 0
    var_epoch_saved = None
    scheduler = None
    if var_mode == "multi_head":                        # train.py:57-63: per-step cosine schedule with linear warm-up
        from .preset import preset
        sch = preset["nn"]["scheduler"]
        scheduler = cosine_schedule_with_warmup(optimizer, sch["num_warmup_epochs"] * total_batches,
                                                preset["nn"]["epoch"] * total_batches, sch["min_lr_ratio"])
    for var_epoch in range(var_epochs):
        var_time_e0 = time.time()
        model.train()
        if sampler is not None:
            sampler.set_epoch(var_epoch)
        predict_train_y = data_batch_y = var_loss_train = None
        if fused:
            # the LAST batch of the epoch is skipped (train.py:81-82): it is never even copied
            lists = _epoch_index_batches(data_train_set, var_batch_size, sampler)[:-1]
            for batch in source.batches(lists):
                data_batch_y = batch.y
                if var_mode == "count_classification":
                    data_batch_y = data_batch_y.sum(axis=1)                              # train.py:91-92
                if var_mode == "baseline":
                    data_batch_y = data_batch_y.reshape(data_batch_y.shape[0], -1)
                var_loss_train, predict_train_y = model.fused_train_step(
                    batch.x, data_batch_y, optimizer, pos_weight=pos_weight, augment=True, grad_hook=sync,
                    loss_kind=loss_kind, offs=batch.offs, lens=batch.lens)
                if scheduler is not None:
                    scheduler.step()
            if predict_train_y is not None:
                var_loss_train, predict_train_y = var_loss_train.clone(), predict_train_y.clone()
                data_batch_y = data_batch_y.clone()
        else:
            for batch_idx, data_batch in enumerate(data_train_loader):
                if batch_idx == total_batches - 1:
                    continue
                data_batch_x, data_batch_y = data_batch
                data_batch_x = data_batch_x.to(device, non_blocking=True)
                data_batch_y = data_batch_y.to(device, non_blocking=True)
                if var_mode == "count_classification":
                    data_batch_y = data_batch_y.sum(axis=1)                              # train.py:91-92
                if var_mode == "baseline":
                    data_batch_y = data_batch_y.reshape(data_batch_y.shape[0], -1)
                if model.training:
                    data_batch_x = apply_augmentation(data_batch_x)
                predict_train_y = model(data_batch_x)
                var_loss_train = loss(predict_train_y, data_batch_y.float())
                optimizer.zero_grad()
                var_loss_train.backward()
                if sync is not None:
                    sync.hook(model._engine)
                optimizer.step()
                if scheduler is not None:
                    scheduler.step()                                                     # train.py:101-102
        if predict_train_y is None:
            raise ValueError("training set smaller than two batches: the reference loop skips the last batch (train.py:81)")
        data_batch_y = data_batch_y.detach().cpu().numpy()
        predict_train_y = predict_train_y.detach().cpu().numpy()
        dict_error_train = performance_metrics(data_batch_y.astype(int), predict_train_y.astype(int), var_mode=var_mode,
                                               var_threshold=var_threshold)
        model.eval()
        if distributed and isinstance(model, ArenaModule):
            broadcast_buffers(model)                    # rank 0's BatchNorm running statistics: identical eval on every rank
        with torch.no_grad():
            data_test_x, data_test_y = next(iter(data_test_loader))
            data_test_x, data_test_y = data_test_x.to(device), data_test_y.to(device)
            if var_mode == "count_classification":
                data_test_y = data_test_y.sum(axis=1)                                    # train.py:116-117
            if var_mode == "baseline":
                data_test_y = data_test_y.reshape(data_test_y.shape[0], -1)
            predict_test_y = model(data_test_x)
            var_loss_test = loss(predict_test_y, data_test_y.float())
            data_test_y = data_test_y.detach().cpu().numpy()
            predict_test_y = predict_test_y.detach().cpu().numpy()
            dict_error_test = performance_metrics(data_test_y, predict_test_y, var_mode, var_threshold)
        _log({
            "epoch": var_epoch, "train_loss": var_loss_train.item(), "test_loss": var_loss_test.item(),
            "total_error_train": dict_error_train["total_error"], "total_error_test": dict_error_test["total_error"],
            "perfect_prediction_percentage_test": dict_error_test["perfect_prediction_percentage"],
            "perfect_prediction_percentage_train": dict_error_train["perfect_prediction_percentage"],
            "accuracy_test": dict_error_test["accuracy"], "accuracy_train": dict_error_train["accuracy"],
            "learning_rate": optimizer.param_groups[0]["lr"], "precision": dict_error_test["precision"],
            "recall": dict_error_test["recall"], "f1_score": dict_error_test["f1_score"],
        })
        print(f"Epoch {var_epoch}/{var_epochs}", "- %.6fs" % (time.time() - var_time_e0),
              "- Loss %.6f" % float(var_loss_train), "- Test Loss %.6f" % float(var_loss_test),
              "- Total Error %.6f" % dict_error_test["total_error"],
              "- Perfect Prediction Percentage Train %.6f" % dict_error_train["perfect_prediction_percentage"],
              "- Perfect Prediction Percentage Test %.6f" % dict_error_test["perfect_prediction_percentage"],
              "- Accuracy Test %.6f" % dict_error_test["accuracy"], "- Accuracy Train %.6f" % dict_error_train["accuracy"],
              "- Precision %.6f" % dict_error_test["precision"], "- Recall %.6f" % dict_error_test["recall"],
              "- F1 Score %.6f" % dict_error_test["f1_score"])
        if (dict_error_test["f1_score"] > var_best_f1_score
                and dict_error_test["perfect_prediction_percentage"] > var_best_PPP):
            var_best_PPP = dict_error_test["perfect_prediction_percentage"]
            var_best_f1_score = dict_error_test["f1_score"]
            var_best_weight = deepcopy(model.state_dict())
            var_epoch_saved = var_epoch
            counter = 0
        else:
            counter += 1
        if counter >= patience:
            print(f"Early stopping triggered at epoch {var_epoch}")
            break
    if source is not None:
        source.close()
    if var_best_weight is None:
        # the reference raises UnboundLocalError here (train.py:175); keep the last weights instead of crashing
        var_best_weight = deepcopy(model.state_dict())
    print(f"Epoch that the model was saved {var_epoch_saved}")
    return var_best_weight
