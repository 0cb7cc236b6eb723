"""Tap-sharing wgrad GEMM (gemm_tn3): correctness vs torch and timing vs the per-tap kernel."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps

ops = NativeOps(torch.device("cuda", 0))
lib = ops.lib
GUARD = 16


def make(M, Na, Dp, nlen, k):
    torch.manual_seed(1)
    A = (torch.randn(M, (Na + 15) // 16 * 16, device="cuda") * 0.1).to(torch.bfloat16)
    full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
    X = full[GUARD:GUARD + M]
    Cm = torch.zeros(Na, nlen, k, device="cuda")
    pl = (k - 1) // 2
    segs = [(j - pl, 0, j, nlen) for j in range(k)]
    return A, full, X, Cm, segs, pl


def ref(A, full, M, Na, nlen, k, pl):
    out = torch.zeros(Na, nlen, k, device="cuda")
    a = A[:, :Na].float()
    for j in range(k):
        x = full[GUARD + j - pl: GUARD + j - pl + M, :nlen].float()
        out[:, :, j] = a.t() @ x
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


for (M, Na, Dp, nlen, k) in [(4096, 270, 272, 270, 5), (3000, 150, 160, 150, 2), (4096, 128, 272, 270, 16), (4096, 270, 272, 270, 1)]:
    A, full, X, Cm, segs, pl = make(M, Na, Dp, nlen, k)
    r = ref(A, full, M, Na, nlen, k, pl)
    Cm.zero_()
    ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs)
    torch.cuda.synchronize()
    print(f"M={M} Na={Na} nlen={nlen} k={k} tn3: rel err {((Cm - r).norm() / r.norm()).item():.3e}", flush=True)

for (M, Na, Dp, nlen, k) in [(39424, 270, 272, 270, 1), (39424, 270, 272, 270, 3), (39424, 270, 272, 270, 5), (39424, 960, 272, 270, 1),
                             (39424, 128, 272, 270, 8), (39424, 128, 272, 270, 16), (70144, 150, 160, 150, 3), (70144, 480, 160, 150, 1)]:
    A, full, X, Cm, segs, pl = make(M, Na, Dp, nlen, k)
    fl = 2.0 * M * Na * nlen * k
    res = []
    ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
    res.append(f"tn3 {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s")
    for dbg in (1, 2):
        lib.csi_set_tn_debug(dbg)
        ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
        res.append(f"dbg{dbg} {ms*1e3:7.1f} us")
    lib.csi_set_tn_debug(0)
    print(f"M={M} Na={Na} nlen={nlen} k={k}: " + " | ".join(res), flush=True)
