"""A few launches of attn_fwd / attn_bwd on a THAT stream shape (for ncu): argv = B L d H reps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
B, L, d, H, reps = [int(a) for a in sys.argv[1:6]]
ops = NativeOps(torch.device("cuda", 0))
HALO, GUARD = 2, 16
hd = d // H
hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
rows = B * (L + 2 * HALO)
def buf(ld, dt=torch.bfloat16, fill=True):
    full = (torch.randn(rows + 2 * GUARD, ld, device="cuda") * (0.5 if fill else 0)).to(dt)
    return full[GUARD:GUARD + rows]
qkv, o, do, dqkv = buf(3 * H * hp), buf(H * hp, fill=False), buf(H * hp), buf(3 * H * hp, fill=False)
lse = torch.zeros(B * H * L, device="cuda")
for _ in range(reps):
    ops.attn_fwd(qkv, o, lse, B, L, d, H, hp, HALO)
    ops.attn_bwd(qkv, o, do, dqkv, lse, B, L, d, H, hp, HALO)
torch.cuda.synchronize()
print("ok")
