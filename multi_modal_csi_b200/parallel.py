"""Batch-sharded data parallelism: one process per GPU, replicated weights, gradients summed over ranks.

The reference is single-process (no collective anywhere); this is the one exchange the
data-parallel path adds.  Gradients already live in ONE flat fp32 arena (multi_modal_csi_b200.that.THAT.flat_grads),
so the exchange is a handful of large contiguous ``ncclAllReduce`` calls issued through ``torch.distributed``.
BatchNorm statistics stay per rank (DistributedDataParallel semantics).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    """Averages the flat gradient arena over the ranks between backward and the optimizer step."""

    def __init__(self, model, world_size: int, num_buckets: int = 1):
        self.model = model
        self.world = world_size
        self.num_buckets = max(1, num_buckets)

    def hook(self, engine):
        g = engine.grads
        if self.world <= 1:
            return
        n = g.numel()
        step = (n + self.num_buckets - 1) // self.num_buckets
        for i in range(0, n, step):
            dist.all_reduce(g[i:i + step], op=dist.ReduceOp.AVG if g.is_cuda else dist.ReduceOp.SUM)
        if not g.is_cuda:
            g.div_(self.world)


def broadcast_parameters(model, src: int = 0):
    """Make every rank start from rank ``src``'s weights and BatchNorm buffers."""
    dist.broadcast(model.flat_params, src)
    for b in model.buffers():
        dist.broadcast(b, src)
