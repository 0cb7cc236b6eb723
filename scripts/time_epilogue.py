"""Out-proj shaped NT GEMM (M=39424, N=270, K=320): cost of the epilogue variants (fp32 out / bias / residual / dropout)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
ops = NativeOps(torch.device("cuda", 0))
M, N, Dp = 39424, 270, 320
A = torch.randn(M, Dp, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, Dp, device="cuda") / math.sqrt(Dp)).to(torch.bfloat16)
segs = [(0, 0, 0, Dp)]
rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
bias = torch.randn(N, device="cuda")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, od, b, r, dp in [("bf16", torch.bfloat16, None, False, 0.0), ("fp32", torch.float32, None, False, 0.0),
                           ("fp32+bias", torch.float32, bias, False, 0.0), ("fp32+bias+drop", torch.float32, bias, False, 0.1),
                           ("fp32+bias+res", torch.float32, bias, True, 0.0), ("fp32+bias+res+drop", torch.float32, bias, True, 0.1)]:
    Cm = torch.zeros(M, 272, dtype=od, device="cuda")
    res = torch.randn(M, 272, device="cuda") if r else None
    print(f"{name:20s} {timeit(lambda: ops.gemm_nt(A, W, Cm, M, N, segs, b, res, dp, 3, rng if dp else None)):7.1f} us", flush=True)
