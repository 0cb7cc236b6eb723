"""Times csi_gemm_nt (tcgen05 v3) on the THAT step shapes with the output split into 1..4 column tiles per row tile
(csi_set_gemm_ntn; values below the minimum that fits one 256-column accumulator are raised to it, so N=270 runs with 2
column tiles for ntn=1 and ntn=2).  Result at B=256: more column tiles never helped.  Prints us per call."""
import sys, os, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps

ops = NativeOps(torch.device("cuda", 0))
GUARD = 16
rng = torch.tensor([1, 0], dtype=torch.int64, device="cuda")


def run(M, N, Dp, k, cdt, res, ntn, reps=20):
    ops.lib.csi_set_gemm_ntn(ctypes.c_int(ntn))
    torch.manual_seed(1)                                   # same operands for every ntn: the outputs must agree
    full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
    A = full[GUARD:GUARD + M]
    W = (torch.randn(N, k * Dp, device="cuda") / math.sqrt(k * Dp)).to(torch.bfloat16)
    ldc = (N + 15) // 16 * 16
    C = torch.zeros(M, ldc, dtype=cdt, device="cuda")
    R = torch.randn(M, ldc, device="cuda") if res else None
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
    for _ in range(3):
        ops.gemm_nt(A, W, C, M, N, segs, None, R, 0.1 if res else 0.0, 3, rng)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm_nt(A, W, C, M, N, segs, None, R, 0.1 if res else 0.0, 3, rng)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, C.float()


bf, f32 = torch.bfloat16, torch.float32
shapes = [("qkv", 39424, 960, 272, 1, bf, False), ("outproj", 39424, 270, 320, 1, f32, True),
          ("conv1", 39424, 270, 272, 1, bf, False), ("conv3", 39424, 270, 272, 3, bf, False),
          ("conv5", 39424, 270, 272, 5, bf, False), ("dgrad9", 39424, 270, 272, 9, bf, False),
          ("head8", 39424, 128, 272, 8, bf, False), ("head16", 39424, 128, 272, 16, bf, False),
          ("hdgrad", 39424, 270, 128, 24, bf, False), ("qkv_dg", 39424, 270, 960, 1, bf, False),
          ("o_dg", 39424, 320, 272, 1, bf, False),
          ("r_qkv", 70144, 480, 160, 1, bf, False), ("r_out", 70144, 150, 160, 1, f32, True),
          ("r_conv3", 70144, 150, 160, 3, bf, False), ("r_dg6", 70144, 150, 160, 6, bf, False),
          ("l540_conv5", 39424, 540, 544, 5, bf, False), ("l540_out", 39424, 540, 640, 1, f32, True)]
for name, M, N, Dp, k, cdt, res in shapes:
    out, ref = [], None
    for ntn in (1, 2, 3, 4):
        us, C = run(M, N, Dp, k, cdt, res, ntn)
        if ref is None:
            ref = C
        err = ((C - ref).norm() / ref.norm()).item()
        out.append(f"ntn={ntn}: {us:7.1f} us" + ("" if err < 1e-6 else f" (DIFF {err:.2e})"))
    print(f"{name:10s} M={M} N={N} K={k}x{Dp}  " + "  ".join(out), flush=True)
ops.lib.csi_set_gemm_ntn(ctypes.c_int(0))
