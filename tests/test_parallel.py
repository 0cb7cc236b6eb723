"""Data-parallel host logic on CPU: world_size-2 gloo run of GradSync + the sharded train loop (mirror kernels)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WANDB_MODE"] = "disabled"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mirror_ops import MirrorOps
    from multi_modal_csi_b200 import THAT, FusedAdam
    from multi_modal_csi_b200.parallel import GradSync, broadcast_parameters
    T, F, out, B = 400, 30, 12, 2
    torch.manual_seed(39 + rank)                       # different init per rank: broadcast must fix it
    m = THAT((T, F), (out,), act_dtype="fp32")
    m._ops_override = MirrorOps()
    m.dropout_enabled = False
    broadcast_parameters(m, 0)
    g = torch.Generator().manual_seed(100 + rank)      # each rank its own shard
    x = torch.rand(B, T, F, generator=g) * 20
    y = (torch.rand(B, out, generator=g) < 0.2).float()
    m.train()
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    sync = GradSync(m, world)
    captured = {"local": torch.zeros_like(m.flat_grads)}

    class Spy:                                       # records the local gradients of each bucket before it is reduced
        def start_bucket(self, eng, lo, hi):
            captured["local"][lo:hi] = eng.grads[lo:hi]
            sync.start_bucket(eng, lo, hi)

        def finish(self, eng, lo, hi):
            captured["local"][lo:hi] = eng.grads[lo:hi]
            sync.finish(eng, lo, hi)
            captured["synced"] = eng.grads.clone()
    m.fused_train_step(x, y, opt, augment=False, grad_hook=Spy())
    torch.save({"local": captured["local"], "synced": captured["synced"], "params": m.flat_params.clone()},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    mean = (r0["local"] + r1["local"]) / 2
    assert not torch.allclose(r0["local"], r1["local"])
    assert torch.allclose(r0["synced"], mean, rtol=1e-5, atol=1e-7)
    assert torch.equal(r0["synced"], r1["synced"])
    assert torch.equal(r0["params"], r1["params"])          # same start + same averaged gradient -> same weights
