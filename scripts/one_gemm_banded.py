"""A few launches of the fused three-branch conv GEMM (csi_gemm_nt_banded) at the left-stream shape of B=256, F=270 (for ncu):
argv = [M d reps]."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
M = int(sys.argv[1]) if len(sys.argv) > 1 else 39424
d = int(sys.argv[2]) if len(sys.argv) > 2 else 270
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kernels = (1, 3, 5)
ops = NativeOps(torch.device("cuda", 0))
GUARD, Dp = 16, (d + 15) // 16 * 16
full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
A = full[GUARD:GUARD + M]
shifts = sorted({t - (k - 1) // 2 for k in kernels for t in range(k)})
W = torch.zeros(3 * Dp, len(shifts) * Dp, dtype=torch.bfloat16, device="cuda")
for j, k in enumerate(kernels):
    for t in range(k):
        c = shifts.index(t - (k - 1) // 2)
        W[j * Dp:j * Dp + d, c * Dp:c * Dp + d] = (torch.randn(d, d, device="cuda") / math.sqrt(d * k)).to(torch.bfloat16)
segs = [(sh, 0, t * Dp, Dp) for t, sh in enumerate(shifts)]
bands = []
for sh in shifts:
    has = [j for j, k in enumerate(kernels) if -((k - 1) // 2) <= sh <= k - 1 - (k - 1) // 2]
    bands.append((has[0] * Dp, (has[-1] + 1) * Dp))
z = torch.zeros(M, 3 * Dp, dtype=torch.bfloat16, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.gemm_nt_banded(A, W, z, M, 3 * Dp, segs, bands, None, None, 0.0, 0, None)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ops.gemm_nt_banded(A, W, z, M, 3 * Dp, segs, bands, None, None, 0.0, 0, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * M * d * d * sum(kernels)
print(f"banded conv trio M={M} d={d}: {ms * 1e3:.1f} us per launch, {fl / ms / 1e9:.0f} TF/s algorithmic")
