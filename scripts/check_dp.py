"""Data-parallel consistency on real GPUs (run under torchrun, one rank per GPU): after a few fused train steps with
different data per rank, every rank must hold bit-identical parameters (the two all-reduced buckets cover the whole
gradient arena and every rank applies the same averaged gradient), and the bucketed/overlapped step must agree with the
plain "one all-reduce after backward" hook."""
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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
