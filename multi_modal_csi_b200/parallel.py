"""Batch-sharded data parallelism: one process per GPU, replicated weights, gradients averaged over ranks.

The reference is single-process (no collective anywhere); this is the one exchange the
data-parallel path adds.  Gradients already live in ONE flat fp32 arena (multi_modal_csi_b200.that.THAT.flat_grads),
so the exchange is two large contiguous ``ncclAllReduce`` calls issued through ``torch.distributed``:

  bucket 1 = [bucket_split, n)   left encoders 1.. + left head, right stream, output layer (~3/4 of the bytes), final
                                 once backward part 1 is done -> reduced on a side stream WHILE left encoder 0's
                                 backward (part 2) runs
  bucket 2 = [0, bucket_split)   Gaussian encoding + left encoder 0, reduced after backward
(``THATEngine.buckets``; the two THAT streams themselves run side by side on two CUDA streams, ``THATEngine._fork``)

BatchNorm statistics stay per rank (DistributedDataParallel semantics).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    """Averages the flat gradient arena over the ranks between backward and the optimizer step.

    Usable as a plain hook (``sync(engine)``: one all-reduce after backward) or through ``start_bucket`` /
    ``finish`` (what ``THAT.fused_train_step`` calls when it sees them) to overlap with backward."""

    def __init__(self, model, world_size: int):
        self.model = model
        self.world = world_size
        self._side = None
        self._work = None
        # the whole data-parallel step as ONE CUDA graph with the all-reduces captured inside (CSI_DP_ONE_GRAPH=0: two graphs with
        # the first all-reduce issued between them)
        import os
        self.one_graph = os.environ.get("CSI_DP_ONE_GRAPH", "1") != "0"

    def _reduce(self, t, async_op=False):
        if t.is_cuda:
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)      # gloo (CPU tests): no AVG
        t.div_(self.world)
        return w

    def hook(self, engine):
        if self.world > 1:
            self._reduce(engine.grads)

    __call__ = hook

    def start_bucket(self, engine, lo: int, hi: int, extra_streams=()):
        """All-reduce grads[lo:hi] on the side stream once the current stream AND every stream in ``extra_streams`` have
        reached this point; the current stream itself does not wait (``finish`` does)."""
        if self.world <= 1:
            return
        g = engine.grads[lo:hi]
        if not g.is_cuda:
            self._reduce(g)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(g.device)
        self._side.wait_stream(torch.cuda.current_stream(g.device))
        for st in extra_streams:
            self._side.wait_stream(st)
        with torch.cuda.stream(self._side):
            self._work = self._reduce(g, async_op=True)

    def finish(self, engine, lo: int, hi: int):
        if self.world <= 1:
            return
        self._reduce(engine.grads[lo:hi])
        if self._work is not None:
            self._work.wait()                      # the compute stream waits for bucket 1's all-reduce
            self._work = None
            torch.cuda.current_stream(engine.grads.device).wait_stream(self._side)      # (rejoins the side stream when captured)


def bind_to_gpu_numa_node(device) -> dict:
    """Pin this process (one rank per GPU) to the CPU cores that are local to its GPU's PCIe root, BEFORE any pinned host
    memory is allocated / registered.

    The input path moves 0.83-1.7 GB per step and GPU over PCIe (train.py:84-86).  Page-locked memory is placed on the NUMA
    node of the thread that first touches it; under torchrun every rank starts on an arbitrary core, so on a two-socket
    host half of the ranks stream their batches across the socket interconnect and all of them share one node's memory
    controllers -- the round-1 scaling runs saw the aggregate H2D rate stop at ~110 GB/s for 2 AND 4 GPUs.  Binding is a
    no-op (returns {"bound": False, ...}) when sysfs does not expose the topology, e.g. in a container with one node."""
    import os
    info = {"bound": False}
    try:
        dev = torch.device(device)
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info.update({"pci": bdf, "numa_node": node, "local_cpus": len(cpus), "usable": len(use)})
        if node >= 0 and use and use != allowed:
            os.sched_setaffinity(0, use)
            torch.set_num_threads(max(1, min(torch.get_num_threads(), len(use))))
            info["bound"] = True
    except Exception as e:                               # topology not visible: leave the scheduler alone
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def broadcast_buffers(model, src: int = 0):
    """Rank ``src``'s BatchNorm running statistics on every rank (they are per rank during training: DDP semantics).
    ``train()`` calls this before each evaluation so that metrics, best-weight selection and early stopping agree on all
    ranks; one flattened broadcast instead of one per buffer."""
    bufs = [b for b in model.buffers()]
    if not bufs:
        return
    flat = torch.cat([b.detach().reshape(-1).to(torch.float64) for b in bufs])      # float64 holds the int64 batch counters exactly
    dist.broadcast(flat, src)
    pos = 0
    for b in bufs:
        n = b.numel()
        b.copy_(flat[pos:pos + n].view(b.shape).to(b.dtype))
        pos += n


def broadcast_parameters(model, src: int = 0):
    """Make every rank start from rank ``src``'s weights and BatchNorm buffers."""
    dist.broadcast(model.flat_params, src)
    broadcast_buffers(model, src)
