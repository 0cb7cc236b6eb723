"""Batch feeding for ``train()``: the B200-side replacement of ``DataLoader(TensorDataset, pin_memory=True)`` +
``.to(device)`` (benchmark/wifi_csi/train.py:48,84-86) and of the padded dataset tensor (load_data.py:62-78).

The reference collates every batch on the host (index + stack: one 829 MB copy at B=256), pins it (a second copy) and
then moves it over PCIe.  Here a batch is never materialised on the host:

* ``resident`` mode -- the whole dataset lives in HBM as ONE fp32 arena (uploaded once; 180 GB of HBM3e holds ~50 000
  padded recordings) and a batch is just ``(arena, offs[B], lens[B])``: ``csi_pool_dual`` gathers, front-pads and pools
  the samples straight out of the arena.  Per step only the offsets, lengths and labels cross PCIe (a few KB).
* ``stream`` mode -- for datasets larger than the HBM budget: the dataset stays in page-locked host memory (registered in
  place, no copy) and batch i+1 is copied sample by sample, straight from the dataset rows, into one of two device
  staging arenas on a copy stream while batch i trains (double buffering, events both ways).

Both modes take dense ``[N, T, ...]`` tensors (offs = i*T*F, lens = T) and the ragged arena of
``load_data.load_data_x_packed`` (unpadded recordings: ~4.5 % fewer bytes, front pad applied inside the pooling kernel).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset


class PackedCSIDataset(Dataset):
    """Unpadded recordings in one fp32 arena + (offset, length) tables (``load_data.load_data_x_packed``) with their
    labels.  ``train()`` feeds it to the kernels without padding; ``__getitem__`` returns the FRONT-padded [T, F] sample
    (load_data.py:66-72) so that any other consumer sees what the reference's TensorDataset would give."""

    def __init__(self, arena, offs, lens, F: int, y, T: int):
        self.arena = torch.as_tensor(arena, dtype=torch.float32).reshape(-1)
        self.offs = torch.as_tensor(np.asarray(offs), dtype=torch.int64)
        self.lens = torch.as_tensor(np.asarray(lens), dtype=torch.int32)
        self.F, self.T = int(F), int(T)
        self.y = torch.as_tensor(y)
        if not (len(self.offs) == len(self.lens) == len(self.y)):
            raise ValueError("offs, lens and y must have one entry per recording")
        if len(self.lens) and int(self.lens.max()) > self.T:
            raise ValueError("recording longer than T")                     # np.pad raises in the reference (load_data.py:70)

    def __len__(self):
        return len(self.lens)

    def __getitem__(self, i):
        n, o = int(self.lens[i]), int(self.offs[i])
        x = torch.zeros(self.T, self.F)
        x[self.T - n:] = self.arena[o:o + n * self.F].view(n, self.F)
        return x, self.y[i]


@dataclass
class DeviceBatch:
    """What ``THAT.fused_train_step(x, y, ..., offs=, lens=)`` takes."""
    x: torch.Tensor                    # fp32 device arena (resident dataset or staging buffer)
    offs: torch.Tensor                 # int64 [B] element offsets into x
    lens: torch.Tensor                 # int32 [B] rows per sample (<= T)
    y: torch.Tensor                    # labels [B, ...] on the device
    size: int
    h2d_bytes: int                     # bytes this batch moved over PCIe


def _host_register(t: torch.Tensor) -> bool:
    """Page-lock an existing host tensor in place (cudaHostRegister): async copies out of it, no second host copy."""
    if t.is_pinned() or t.numel() == 0:
        return t.is_pinned()
    try:
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        return int(rc) == 0
    except Exception:
        return False


def _host_unregister(t: torch.Tensor):
    try:
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
    except Exception:
        pass


class CSIBatchSource:
    def __init__(self, dataset, device, batch_size: int, mode: str = "auto", hbm_fraction: float = 0.5):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CSIBatchSource feeds the CUDA path; there is no CPU path")
        self.B = int(batch_size)
        if isinstance(dataset, PackedCSIDataset):
            self.x_host, self.offs_host, self.lens_host = dataset.arena, dataset.offs, dataset.lens
            self.y_host, self.T, self.F = dataset.y, dataset.T, dataset.F
        else:                                                   # TensorDataset(x [N,T,...], y): dense, offs = i*T*F
            x, y = dataset.tensors[0], dataset.tensors[1]
            if x.dtype != torch.float32:
                x = x.float()
            x = x.contiguous()
            N, T = x.shape[0], x.shape[1]
            self.F = int(x[0].numel() // T) if N else 0
            self.T = int(T)
            self.x_host = x.reshape(-1)
            self.offs_host = torch.arange(N, dtype=torch.int64) * (self.T * self.F)
            self.lens_host = torch.full((N,), self.T, dtype=torch.int32)
            self.y_host = y
        self.N = len(self.lens_host)
        nbytes = self.x_host.numel() * 4
        if mode == "auto":
            free, _total = torch.cuda.mem_get_info(self.device)
            mode = "resident" if nbytes <= hbm_fraction * free else "stream"
        if mode not in ("resident", "stream"):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode
        self.copy_stream = torch.cuda.Stream(self.device)
        self._registered = False
        self.upload_bytes = 0
        y_row = self.y_host[0].numel() if self.N else 0
        # small per-batch tables (offsets, lengths, labels): pinned staging, two slots
        self._tab = [dict(offs=torch.empty(self.B, dtype=torch.int64).pin_memory(),
                          lens=torch.empty(self.B, dtype=torch.int32).pin_memory(),
                          y=torch.empty((self.B,) + tuple(self.y_host.shape[1:]), dtype=self.y_host.dtype).pin_memory(),
                          d_offs=torch.empty(self.B, dtype=torch.int64, device=self.device),
                          d_lens=torch.empty(self.B, dtype=torch.int32, device=self.device),
                          d_y=torch.empty((self.B,) + tuple(self.y_host.shape[1:]), dtype=self.y_host.dtype, device=self.device),
                          ready=torch.cuda.Event(), freed=torch.cuda.Event()) for _ in range(2)]
        self._y_row = y_row
        if mode == "resident":
            self._upload()
        else:
            self._registered = _host_register(self.x_host)
            if not self._registered and not self.x_host.is_pinned():
                self.x_host = self.x_host.pin_memory()          # fallback: one pinned copy of the dataset
            cap = self.B * self.T * self.F
            self._stage = [torch.empty(cap, dtype=torch.float32, device=self.device) for _ in range(2)]

    def close(self):
        if self._registered:
            torch.cuda.synchronize(self.device)
            _host_unregister(self.x_host)
            self._registered = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ resident mode: one upload
    def _upload(self, chunk_bytes: int = 256 << 20):
        n = self.x_host.numel()
        self.x_dev = torch.empty(n, dtype=torch.float32, device=self.device)
        step = chunk_bytes // 4
        bounce = [torch.empty(min(step, max(n, 1)), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        with torch.cuda.stream(self.copy_stream):
            for k, lo in enumerate(range(0, n, step)):
                hi, s = min(n, lo + step), k % 2
                if k >= 2:
                    done[s].synchronize()                       # the bounce buffer's previous copy has left the host
                bounce[s][:hi - lo].copy_(self.x_host[lo:hi])
                self.x_dev[lo:hi].copy_(bounce[s][:hi - lo], non_blocking=True)
                done[s].record(self.copy_stream)
        self.copy_stream.synchronize()
        self.upload_bytes = n * 4

    # ------------------------------------------------------------------ iteration
    def _issue(self, idx: Sequence[int], slot: int) -> DeviceBatch:
        """Queue the copies of one batch on the copy stream; returns the device-side batch (valid after ``ready``)."""
        tab = self._tab[slot]
        b = len(idx)
        it = torch.as_tensor(idx, dtype=torch.int64)
        lens = self.lens_host[it]
        tab["lens"][:b] = lens
        tab["y"][:b] = self.y_host[it]
        nbytes = b * (8 + 4) + b * self._y_row * self.y_host.element_size()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(tab["freed"])           # the step that used this slot has consumed it
            if self.mode == "resident":
                tab["offs"][:b] = self.offs_host[it]
                x = self.x_dev
            else:
                x = self._stage[slot]
                pos = 0
                src_offs = self.offs_host[it].tolist()
                ll = lens.tolist()
                # one async copy per RUN of samples that are adjacent in the source arena (a shuffled batch: one per
                # sample; an unshuffled / sorted one: a few large copies)
                run_src = run_dst = run_len = 0
                for i in range(b):
                    n = ll[i] * self.F
                    tab["offs"][i] = pos
                    if run_len and src_offs[i] == run_src + run_len:
                        run_len += n
                    else:
                        if run_len:
                            x[run_dst:run_dst + run_len].copy_(self.x_host[run_src:run_src + run_len], non_blocking=True)
                        run_src, run_dst, run_len = src_offs[i], pos, n
                    pos += n
                if run_len:
                    x[run_dst:run_dst + run_len].copy_(self.x_host[run_src:run_src + run_len], non_blocking=True)
                nbytes += pos * 4
            tab["d_offs"][:b].copy_(tab["offs"][:b], non_blocking=True)
            tab["d_lens"][:b].copy_(tab["lens"][:b], non_blocking=True)
            tab["d_y"][:b].copy_(tab["y"][:b], non_blocking=True)
            tab["ready"].record(self.copy_stream)
        return DeviceBatch(x, tab["d_offs"][:b], tab["d_lens"][:b], tab["d_y"][:b], b, nbytes)

    def batches(self, index_batches: Iterable[Sequence[int]]):
        """Yields a DeviceBatch per index list; the copies of the next batch are in flight while the caller trains on
        the current one.  The caller's work on a batch must be queued on the current stream before it asks for the next."""
        lists: List[Sequence[int]] = [list(b) for b in index_batches]
        if not lists:
            return
        cur = torch.cuda.current_stream(self.device)
        for s in range(2):
            self._tab[s]["freed"].record(cur)
        # the pinned tables of a slot are rewritten by the host: wait until their previous H2D copy has been issued AND done
        nxt = self._issue(lists[0], 0)
        for k in range(len(lists)):
            slot = k % 2
            batch = nxt
            if k + 1 < len(lists):
                self._tab[1 - slot]["ready"].synchronize()      # slot's previous table copies have completed: host may rewrite them
                nxt = self._issue(lists[k + 1], 1 - slot)
            cur.wait_event(self._tab[slot]["ready"])
            yield batch
            self._tab[slot]["freed"].record(cur)
