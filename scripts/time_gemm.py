"""Times csi_gemm_nt (tcgen05 vs FFMA) on the THAT layer shapes; prints TFLOP/s."""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps

ops = NativeOps(torch.device("cuda", 0))
GUARD = 16

def run(M, N, Dp, k, simt, cdt=torch.bfloat16, reps=5):
    ops.set_force_simt(simt)
    full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
    A = full[GUARD:GUARD + M]
    W = (torch.randn(N, k * Dp, device="cuda") / math.sqrt(k * Dp)).to(torch.bfloat16)
    C = torch.zeros(M, (N + 15) // 16 * 16, dtype=cdt, device="cuda")
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
    t0 = time.time()
    ops.gemm_nt(A, W, C, M, N, segs, None, None, 0.0, 0, None)
    torch.cuda.synchronize()
    first = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm_nt(A, W, C, M, N, segs, None, None, 0.0, 0, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * N * k * Dp
    print(f"M={M} N={N} K={k}x{Dp} simt={simt}: first {first*1e3:.1f} ms, {ms:.3f} ms/call, {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    return C.float()

for (M, N, Dp, k) in [(39424, 810, 272, 1), (39424, 270, 272, 1), (39424, 270, 272, 3), (39424, 270, 272, 5),
                      (39424, 128, 272, 16), (70144, 450, 160, 1), (70144, 150, 160, 3)]:
    a = run(M, N, Dp, k, False)
    b = run(M, N, Dp, k, True, reps=2)
    print("   rel diff", ((a - b).norm() / b.norm()).item(), flush=True)
