"""Per-tensor gradient error of the fp32 kernel path vs the fp64 CPU oracle (and of the fp32 CPU oracle itself)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200 import THAT
from oracle import that_oracle as O

F, out = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
T, B = 3000, 4
torch.manual_seed(39)
m = THAT((T, F), (out,), act_dtype=mode)
m.dropout_enabled = False
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.to("cuda")
g = torch.Generator().manual_seed(1234)
x = torch.rand(B, T, F, generator=g) * 20
y = (torch.rand(B, out, generator=g) < 0.15).float()
m.train()
logits = m(x.cuda())
loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 4.0, device="cuda"))(logits, y.cuda())
loss.backward()
sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
l64, _, g64 = O.loss_and_grads(sd64, x.double(), y.double())
l32, _, g32 = O.loss_and_grads(copy.deepcopy(sd), x, y)
print("logits rel err: ours", ((logits.cpu().double() - l64).norm() / l64.norm()).item(), " cpu-fp32", ((l32.double() - l64).norm() / l64.norm()).item())
rows = []
num = den = num32 = 0.0
for k, p in m.named_parameters():
    if k in g64:
        r = g64[k]
        e = (p.grad.double().cpu() - r).norm().item(); e32 = (g32[k].double() - r).norm().item(); n = r.norm().item()
        num += e * e; den += n * n; num32 += e32 * e32
        rows.append((e, e32, n, k))
print("grad normwise: ours %.3e  cpu-fp32 %.3e  (|g| = %.3f)" % ((num / den) ** 0.5, (num32 / den) ** 0.5, den ** 0.5))
rows.sort(reverse=True)
for e, e32, n, k in rows[:12]:
    print("  abs err %.3e (cpu32 %.3e) norm %.3e  %s" % (e, e32, n, k))
