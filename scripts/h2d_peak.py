"""Concurrent pinned host -> device copy bandwidth on N GPUs of one box (run under torchrun, one rank per GPU), with and
without binding every rank to its GPU's NUMA node: the PCIe / host-memory ceiling `e2e` is bounded by at N > 1.

    python -m torch.distributed.run --nproc-per-node N scripts/h2d_peak.py [MB per copy]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multi_modal_csi_b200.parallel import bind_to_gpu_numa_node

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 829
n = mb * (1 << 20) // 4
dst = torch.empty(n, device=dev)


def measure(tag, info=None):
    src = torch.empty(n).pin_memory()                    # first touched (and placed) by this thread, under its current affinity
    src.fill_(1.0)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 10 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs, gbs], device=dev)
    if world > 1:
        lo = t[:1].clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        sm = t[1:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t = torch.cat([lo, sm])
    if rank == 0:
        print(json.dumps({"what": tag, "n_gpus": world, "mb_per_copy": mb, "min_rank_gbs": round(t[0].item(), 1),
                          "aggregate_gbs": round(t[1].item(), 1), "rank0_binding": info}), flush=True)
    del src


measure("unbound (torchrun default placement)")
info = bind_to_gpu_numa_node(dev)
measure("bound to the GPU's NUMA node", info)
if world > 1:
    dist.destroy_process_group()
