"""Per-parameter gradient errors of CNN_2D (bf16 / fp32) against the oracle at several shapes: argv = B T [F out]."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200 import CNN_2D
from oracle import cnn2d_oracle as C

def nrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

for (B, T) in [(3, 300), (8, 600), (16, 3000)]:
    F, out = 270, 54
    gen = torch.Generator().manual_seed(2468)
    x = torch.rand(B, T, F, generator=gen) * 20
    y = (torch.rand(B, out, generator=gen) < 0.15).float()
    for mode in ("fp32", "bf16"):
        torch.manual_seed(39)
        m = CNN_2D((T, F), (out,), act_dtype=mode)
        sd = copy.deepcopy(m.state_dict())
        m.dropout_enabled = False
        m = m.to("cuda").train()
        logits = m(x.cuda())
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.full((out,), 6.0, device="cuda"))(logits, y.cuda())
        loss.backward()
        rl, rloss, rg = C.loss_and_grads(sd, x, y)
        num = sum((p.grad.double().cpu() - rg[k].double()).pow(2).sum().item() for k, p in m.named_parameters())
        den = sum(rg[k].double().pow(2).sum().item() for k, p in m.named_parameters())
        print(f"B={B} T={T} {mode}: logits {nrel(logits, rl):.3e} loss {loss.item():.6f}/{rloss.item():.6f} grad {(num / den) ** 0.5:.3e}   " +
              " ".join(f"{k.replace('layer_', '')}:{nrel(p.grad, rg[k]):.1e}" for k, p in m.named_parameters()), flush=True)
        if mode == "bf16":
            # how ill-conditioned is the last BatchNorm: |mean| / std per channel of its input
            eng = m._engine
            M = B * m.geom.H[3] * m.geom.W[3]
            y2 = eng.L[2]["y"][:M].float()
            print("    |mean|/std of the final BatchNorm input: median %.2f max %.2f; rows per channel %d" % (
                (y2.mean(0).abs() / y2.std(0)).median().item(), (y2.mean(0).abs() / y2.std(0)).max().item(), M), flush=True)
        del m
        torch.cuda.empty_cache()
