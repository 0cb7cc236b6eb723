"""BASELINE config 5: THAT large-batch sweep B = 64..4096 PER GPU (F = 270, out = 54) on 1 GPU, or on N GPUs under torchrun
(one rank per GPU, batch-sharded data parallel with the overlapped gradient all-reduce).

Per batch size: device-resident train throughput (augment + fwd + BCE + bwd + Adam, dropout on; CUDA events over `steps`
steps after warm-up, max over ranks, whole-job samples/s).  Parity part (dropout / augmentation off, fp32 kernels): the
loss of the first 3 Adam steps against the CPU oracle on the same seeded inputs -- at N > 1 the oracle is run the
data-parallel way (per-shard forward/backward with per-shard BatchNorm statistics, gradients averaged, one Adam step).
One JSON line per batch size on stdout (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import synth_batch, T_LEN
from multi_modal_csi_b200 import THAT, FusedAdam
from multi_modal_csi_b200.parallel import GradSync

F, out = 270, 54
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512, 1024, 2048, 4096]


def parity(B=16 if world > 1 else 64, steps=3):
    """Loss trajectory of `steps` data-parallel Adam steps (B samples per rank) against the oracle run the same way."""
    from oracle import that_oracle as O
    torch.manual_seed(39)
    m = THAT((T_LEN, F), (out,), act_dtype="fp32", max_batch=B)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.dropout_enabled = False
    m = m.to(dev).train()
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(m, world) if world > 1 else None
    names = O.trainable_names(sd)
    st, got, ref = {}, [], []
    for s in range(steps):
        x, y = synth_batch(B, F, out, 1234 + 10 * s + rank)
        loss, _ = m.fused_train_step(x.to(dev), y.to(dev), opt, augment=False, grad_hook=sync)
        t = loss.detach().clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        got.append(float(t.item()))
        if rank == 0:                                   # the oracle, shard by shard
            gsum, lsum = None, 0.0
            for r in range(world):
                xr, yr = synth_batch(B, F, out, 1234 + 10 * s + r)
                leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
                work = dict(sd); work.update(leaves)
                lo = O.bce_with_logits(O.that_forward(work, xr, training=True, drop=None), yr)
                grads = torch.autograd.grad(lo, [leaves[k] for k in names])
                gsum = list(grads) if gsum is None else [a + b for a, b in zip(gsum, grads)]
                lsum += float(lo)
            with torch.no_grad():
                for k, g in zip(names, gsum):
                    if k not in st:
                        st[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
                    O.adam_update(sd[k], g / world, st[k][0], st[k][1], s + 1, 5e-4, 2e-4)
            ref.append(lsum / world)
        if world > 1:
            dist.barrier()
    if rank != 0:
        return None
    return {"parity_B_per_gpu": B, "n_gpus": world, "loss_fused_fp32": got, "loss_oracle_cpu": ref,
            "max_rel_diff": max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(got, ref))}


for B in sizes:
    torch.manual_seed(39)
    m = THAT((T_LEN, F), (out,), act_dtype="bf16", max_batch=B).to(dev).train()
    m.rng_seed = 1000 + rank
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(m, world) if world > 1 else None
    batches = [tuple(t.to(dev) for t in synth_batch(B, F, out, 1234 + i + 100 * rank)) for i in range(2)]
    steps, warm = (20, 4) if B <= 1024 else (8, 3)
    for i in range(warm):
        m.fused_train_step(*batches[i % 2], opt, pos_weight=4.0, augment=True, grad_hook=sync)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss, _ = m.fused_train_step(*batches[i % 2], opt, pos_weight=4.0, augment=True, grad_hook=sync)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({"B_per_gpu": B, "n_gpus": world, "ms_per_step": ms, "samples_per_s": world * B / ms * 1e3,
                          "loss_last": float(loss.item()), "hbm_gb": torch.cuda.max_memory_allocated(dev) / 2**30}), flush=True)
    del m, opt, batches, sync
    torch.cuda.empty_cache()
p = parity()
if rank == 0:
    print(json.dumps(p), flush=True)
if world > 1:
    dist.destroy_process_group()
