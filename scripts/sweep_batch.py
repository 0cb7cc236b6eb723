"""BASELINE config 5: THAT large-batch sweep B = 64..4096 per GPU (F = 270, out = 54) on one GPU.

Per batch size: device-resident train throughput (augment + fwd + BCE + bwd + Adam, dropout on; CUDA events over
`steps` steps after warm-up).  Parity part (dropout / augmentation off, fp32 kernels): the loss of the first 3 Adam steps
at B = 64 against the CPU oracle on the same seeded inputs.  One JSON line per batch size on stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, T_LEN
from multi_modal_csi_b200 import THAT, FusedAdam

F, out = 270, 54
dev = torch.device("cuda", 0)
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512, 1024, 2048, 4096]


def parity(B=64, steps=3):
    from oracle import that_oracle as O
    torch.manual_seed(39)
    m = THAT((T_LEN, F), (out,), act_dtype="fp32", max_batch=B)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.dropout_enabled = False
    m = m.to(dev).train()
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    names = O.trainable_names(sd)
    st, got, ref = {}, [], []
    for s in range(steps):
        x, y = synth_batch(B, F, out, 1234 + s)
        loss, _ = m.fused_train_step(x.to(dev), y.to(dev), opt, augment=False)
        got.append(float(loss.item()))
        leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
        work = dict(sd); work.update(leaves)
        lo = O.bce_with_logits(O.that_forward(work, x, training=True, drop=None), y)
        grads = torch.autograd.grad(lo, [leaves[k] for k in names])
        with torch.no_grad():
            for k, g in zip(names, grads):
                if k not in st:
                    st[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
                O.adam_update(sd[k], g, st[k][0], st[k][1], s + 1, 5e-4, 2e-4)
        ref.append(float(lo))
    return {"parity_B": B, "loss_fused_fp32": got, "loss_oracle_cpu": ref,
            "max_rel_diff": max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(got, ref))}


for B in sizes:
    torch.manual_seed(39)
    m = THAT((T_LEN, F), (out,), act_dtype="bf16", max_batch=B).to(dev).train()
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=2e-4)
    batches = [tuple(t.to(dev) for t in synth_batch(B, F, out, 1234 + i)) for i in range(2)]
    steps, warm = (20, 4) if B <= 1024 else (8, 3)
    for i in range(warm):
        m.fused_train_step(*batches[i % 2], opt, pos_weight=4.0, augment=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss, _ = m.fused_train_step(*batches[i % 2], opt, pos_weight=4.0, augment=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"B": B, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "loss_last": float(loss.item()),
                      "hbm_gb": torch.cuda.max_memory_allocated(dev) / 2**30}), flush=True)
    del m, opt, batches
    torch.cuda.empty_cache()
print(json.dumps(parity()), flush=True)
