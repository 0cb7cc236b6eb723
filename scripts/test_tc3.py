"""Tap-sharing NT GEMM (gemm_tc3): correctness of the row-shifted UMMA descriptor in both base-offset modes vs a torch
reference, and timing against the per-tap kernel (gemm_tc2)."""
import sys, os, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps

ops = NativeOps(torch.device("cuda", 0))
lib = ops.lib
GUARD = 16


def make(M, N, Dp, k):
    torch.manual_seed(1)
    full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
    A = full[GUARD:GUARD + M]
    W = (torch.randn(N, k * Dp, device="cuda") / math.sqrt(k * Dp)).to(torch.bfloat16)
    ldc = (N + 15) // 16 * 16
    Cm = torch.zeros(M, ldc, dtype=torch.bfloat16, device="cuda")
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
    return full, A, W, Cm, segs, pl


def ref(full, W, M, N, Dp, k, pl):
    out = torch.zeros(M, N, device="cuda")
    for j in range(k):
        a = full[GUARD + j - pl: GUARD + j - pl + M].float()
        out += a @ W[:, j * Dp:(j + 1) * Dp].float().t()
    return out


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (M, N, Dp, k) in [(1024, 270, 272, 5), (1000, 150, 160, 2), (2048, 128, 272, 16)]:
    full, A, W, Cm, segs, pl = make(M, N, Dp, k)
    r = ref(full, W, M, N, Dp, k, pl)
    for mode in (0, 1):
        lib.csi_set_gemm_desc_mode(mode)
        Cm.zero_()
        ops.gemm_nt(A, W, Cm, M, N, segs, None, None, 0.0, 0, None)
        torch.cuda.synchronize()
        err = ((Cm[:, :N].float() - r).norm() / r.norm()).item()
        print(f"M={M} N={N} k={k} desc_mode={mode}: rel err {err:.3e}", flush=True)

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.csi_set_gemm_desc_mode(mode)
for (M, N, Dp, k) in [(39424, 960, 272, 1), (39424, 270, 272, 1), (39424, 270, 272, 3), (39424, 270, 272, 5),
                      (39424, 128, 272, 8), (39424, 128, 272, 16), (70144, 480, 160, 1), (70144, 150, 160, 3)]:
    full, A, W, Cm, segs, pl = make(M, N, Dp, k)
    fl = 2.0 * M * N * k * Dp
    res = []
    ms = timeit(lambda: ops.gemm_nt(A, W, Cm, M, N, segs, None, None, 0.0, 0, None))
    res.append(f"tc3 {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s")
    print(f"M={M} N={N} K={k}x{Dp}: " + " | ".join(res), flush=True)
