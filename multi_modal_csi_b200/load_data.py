"""Annotation / CSI-amplitude loading and label encoding with the semantics of benchmark/wifi_csi/load_data.py.

``load_data_x`` keeps the reference's output ([N, T, 3, 3, 30] fp32, FRONT zero-padded); ``load_data_x_packed`` is
the B200-side alternative that keeps recordings unpadded in one arena + (offset, length) tables so that the pad is
applied inside the pooling kernel (csi_pool_dual) instead of being materialised.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from .preset import preset

_USERS = [f"user_{i}" for i in range(1, 7)]


def load_data_y(var_path_data_y, var_environment=None, var_wifi_band=None, var_num_users=None):
    """load_data.py:15-44: annotation rows (all columns as str) filtered by environment / band / #users."""
    df = pd.read_csv(var_path_data_y, dtype=str)
    for col, sel in (("environment", var_environment), ("wifi_band", var_wifi_band), ("number_of_users", var_num_users)):
        if sel is not None:
            df = df[df[col].isin(sel)]
    return df


def load_data_x(var_path_data_x, var_label_list):
    """load_data.py:48-78: np.load each ``<label>.npy`` ([t,3,3,30]) and zero-pad IN FRONT to preset length."""
    T = preset["data"]["length"]
    out = []
    for label in var_label_list:
        a = np.load(os.path.join(var_path_data_x, label + ".npy"))
        if a.shape[0] > T:
            raise ValueError("index can't contain negative values")        # what np.pad raises in the reference
        out.append(np.concatenate([np.zeros((T - a.shape[0],) + a.shape[1:], dtype=a.dtype), a], axis=0))
    return np.array(out)


def load_data_x_packed(var_path_data_x, var_label_list):
    """Unpadded variant: returns (arena fp32 [sum t_i * F], offs int64 [N] in elements, lens int32 [N], F)."""
    T = preset["data"]["length"]
    chunks, lens = [], []
    for label in var_label_list:
        a = np.load(os.path.join(var_path_data_x, label + ".npy")).astype(np.float32, copy=False)
        if a.shape[0] > T:
            raise ValueError("recording longer than preset['data']['length']")
        chunks.append(a.reshape(a.shape[0], -1))
        lens.append(a.shape[0])
    F = chunks[0].shape[1] if chunks else 0
    lens = np.asarray(lens, dtype=np.int32)
    offs = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        offs[1:] = np.cumsum(lens[:-1].astype(np.int64) * F)
    arena = np.concatenate([c.reshape(-1) for c in chunks]) if chunks else np.zeros(0, np.float32)
    return arena, offs, lens, F


def _columns(df, suffix):
    return df[[f"{u}_{suffix}" for u in _USERS]].to_numpy(copy=True).astype(str)


def encode_identity(data_pd_y):
    """load_data.py:111-131: user present <=> its location is annotated.  -> int8 [N, 6]."""
    return (_columns(data_pd_y, "location") != "nan").astype("int8")


def _lookup(values, table):
    return np.array([[table[v] for v in row] for row in values])


def encode_activity(data_pd_y, var_encoding):
    """load_data.py:135-157 -> [N, 6, 9]."""
    return _lookup(_columns(data_pd_y, "activity"), var_encoding)


def encode_location(data_pd_y, var_encoding):
    """load_data.py:161-183 -> [N, 6, 5]."""
    return _lookup(_columns(data_pd_y, "location"), var_encoding)


def encode_data_y(data_pd_y, var_task):
    """load_data.py:82-107."""
    if var_task == "identity":
        return encode_identity(data_pd_y)
    if var_task == "activity":
        return encode_activity(data_pd_y, preset["encoding"]["activity"])
    if var_task == "location":
        return encode_location(data_pd_y, preset["encoding"]["location"])
    return None
