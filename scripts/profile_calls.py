"""Per-call device time of one eager train step (CUDA events around every C-ABI call): time, TFLOP/s, GB/s per launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, out_dim, T_LEN
from multi_modal_csi_b200 import THAT, FusedAdam

F = int(sys.argv[1]) if len(sys.argv) > 1 else 270
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out = out_dim(F)
dev = torch.device("cuda", 0)
torch.manual_seed(39)
model = THAT((T_LEN, F), (out,), act_dtype="bf16", max_batch=B).to(dev)
model.train()
opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=2e-4)
x, y = synth_batch(B, F, out, 1234)
x, y = x.to(dev), y.to(dev)
for _ in range(3):
    model.fused_train_step(x, y, opt, pos_weight=4.0, augment=True, use_graph=False)
ops = model._engine.ops
acc = {}
R = 3
for r in range(R):
    ops.start_profile()
    model.fused_train_step(x, y, opt, pos_weight=4.0, augment=True, use_graph=False)
    ops.stop_profile()
    for i, c in enumerate(ops.last_calls):
        acc.setdefault(i, []).append(c)
tot = 0.0
for i in sorted(acc):
    name, _, fl, nb, tag = acc[i][0]
    ms = min(c[1] for c in acc[i])
    tot += ms
    tf = fl / ms / 1e9 if fl else 0
    gb = nb / ms / 1e6 if nb else 0
    print(f"{i:4d} {name:20s} {ms*1e3:8.1f} us {tf:8.1f} TF/s {gb:8.0f} GB/s  {tag}")
print(f"total {tot:.3f} ms")
