"""Data-parallel consistency on real GPUs (run under torchrun, one rank per GPU; tests/test_gpu_dp.py launches it):

  1. after a few fused train steps with different data per rank every rank holds bit-identical parameters (the two
     all-reduced buckets cover the whole gradient arena and every rank applies the same averaged gradient),
  2. the bucketed / overlapped step agrees with the plain "one all-reduce after backward" hook,
  3. the all-reduced gradient equals the ORACLE's: the mean over ranks of the CPU oracle's gradient on each rank's shard
     (BatchNorm statistics are per rank -- DistributedDataParallel semantics -- so the oracle is evaluated shard by shard),
     fp32 kernels, 1e-4 on the full gradient vector, and the loss of each rank matches the oracle's shard loss.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import synth_batch
from multi_modal_csi_b200 import THAT, FusedAdam
from multi_modal_csi_b200.parallel import GradSync, broadcast_parameters

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, F, out, T = 16, 270, 54, 3000
res = {}
for mode in ("overlap", "plain"):
    torch.manual_seed(39 + rank)                       # different init per rank: broadcast must fix it
    m = THAT((T, F), (out,), act_dtype="bf16", max_batch=B).to(dev)
    m.dropout_enabled = False
    broadcast_parameters(m, 0)
    m.train()
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(m, world)
    hook = sync if mode == "overlap" else sync.hook
    x, y = synth_batch(B, F, out, 500 + rank)
    x, y = x.to(dev), y.to(dev)
    losses = []
    for s in range(4):                                 # step 0 eager, then CUDA-graph replays
        loss, _ = m.fused_train_step(x, y, opt, augment=False, grad_hook=hook)
        losses.append(float(loss))
    p = m.flat_params.clone()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    res[mode] = (p, losses, same)
    if rank == 0:
        print(f"{mode}: ranks identical = {same}, losses {['%.5f' % l for l in losses]}", flush=True)
d = ((res["overlap"][0] - res["plain"][0]).norm() / res["plain"][0].norm()).item()
if rank == 0:
    print(f"overlap vs plain parameters: rel diff {d:.3e}", flush=True)
ok = res["overlap"][2] and res["plain"][2] and d < 1e-3

# ---- 3. one data-parallel step against the oracle (fp32 kernels)
from oracle import that_oracle as O
Bo = 8
torch.manual_seed(39)
m = THAT((T, F), (out,), act_dtype="fp32", max_batch=Bo)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m.dropout_enabled = False
m = m.to(dev).train()
opt = FusedAdam(m.parameters(), lr=0.0, weight_decay=0.0)          # lr 0: the step leaves the synced gradient in the arena
x, y = synth_batch(Bo, F, out, 900 + rank)
loss, _ = m.fused_train_step(x.to(dev), y.to(dev), opt, augment=False, grad_hook=GradSync(m, world))
torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
_, ref_loss, ref_grads = O.loss_and_grads(sd, x, y)
flat_ref = torch.zeros_like(m.flat_grads)
for k, off in m.arena.offsets.items():
    if k in ref_grads:
        flat_ref[off:off + ref_grads[k].numel()] = ref_grads[k].reshape(-1).to(dev)
dist.all_reduce(flat_ref, op=dist.ReduceOp.AVG)                    # the oracle's data-parallel gradient
gerr = ((m.flat_grads - flat_ref).norm() / flat_ref.norm()).item()
lerr = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
t = torch.tensor([gerr, lerr], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"DP gradient vs oracle (mean of per-shard oracle gradients, {world} ranks x B={Bo}, fp32): rel {t[0].item():.3e}; "
          f"shard loss vs oracle: rel {t[1].item():.3e}", flush=True)
ok = ok and t[0].item() < 1e-4 and t[1].item() < 1e-4
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
