"""A few launches of one NT GEMM shape (for ncu): argv = M N Dp k reps [res]  (res: fp32 output + bias + dropout 0.1 + fp32 residual)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
M, N, Dp, k, reps = [int(a) for a in sys.argv[1:6]]
ops = NativeOps(torch.device("cuda", 0))
GUARD = 16
full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
A = full[GUARD:GUARD + M]
W = (torch.randn(N, k * Dp, device="cuda") / math.sqrt(k * Dp)).to(torch.bfloat16)
RES = len(sys.argv) > 6 and sys.argv[6] == "res"
Cm = torch.zeros(M, (N + 15) // 16 * 16, dtype=torch.float32 if RES else torch.bfloat16, device="cuda")
bias = torch.randn(N, device="cuda") if RES else None
res = torch.randn_like(Cm) if RES else None
rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda") if RES else None
pl = (k - 1) // 2
segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
for _ in range(reps):
    ops.gemm_nt(A, W, Cm, M, N, segs, bias, res, 0.1 if RES else 0.0, 3, rng)
torch.cuda.synchronize()
print("ok")
