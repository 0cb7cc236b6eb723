"""Tap-sharing wgrad GEMM (gemm_tn3): correctness vs torch, the two reduction modes (fp32 atomics from every CTA vs partial sums +
fixed-order reduce, csi_gemm_tn_workspace), run-to-run reproducibility and timing."""
import sys, os, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200 import ops as OPS
from multi_modal_csi_b200.ops import NativeOps

ops = NativeOps(torch.device("cuda", 0))
lib = ops.lib
GUARD = 16


def two_stage(on):
    OPS.TN_TWO_STAGE = on
    if not on:
        for (_dev, st) in list(OPS._TN_WS):
            lib.csi_gemm_tn_workspace(C.c_void_p(st), C.c_void_p(0), C.c_longlong(0))
        OPS._TN_WS.clear()


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
    """device time per call, the calls replayed from a CUDA graph (no host launch overhead between them)"""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mode in (False, True):
    two_stage(mode)
    for (M, Na, Dp, nlen, k) in [(4096, 270, 272, 270, 5), (3000, 150, 160, 150, 2), (4096, 128, 272, 270, 16), (4096, 270, 272, 270, 1),
                                 (39424, 270, 272, 270, 3), (70144, 150, 160, 150, 3), (1000, 54, 288, 288, 1)]:
        A, full, X, Cm, segs, pl = make(M, Na, Dp, nlen, k)
        r = ref(A, full, M, Na, nlen, k, pl)
        outs = []
        for _ in range(2):
            Cm.zero_()
            ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs)
            torch.cuda.synchronize()
            outs.append(Cm.clone())
        print(f"two_stage={mode} M={M} Na={Na} nlen={nlen} k={k}: rel err {((Cm - r).norm() / r.norm()).item():.3e}  "
              f"bit-identical reruns: {bool(torch.equal(outs[0], outs[1]))}", flush=True)
    # head-padded operands (in-projection / out-projection weight gradients): compact indices on either side
    H, hd, hp, d, rows = 10, 27, 32, 270, 4096
    A2 = (torch.randn(rows, H * hp, device="cuda") * 0.1).to(torch.bfloat16)
    Xf = torch.randn(rows + 2 * GUARD, 272, device="cuda").to(torch.bfloat16)
    X = Xf[GUARD:GUARD + rows]
    c1 = torch.zeros(H * hd, d, device="cuda")
    ops.gemm_tn(A2, X, c1, d, 1, rows, H * hp, [(0, 0, 0, d)], (hd, hp), (0, 0))
    a2c = A2.float().view(rows, H, hp)[:, :, :hd].reshape(rows, H * hd)
    r1 = a2c.t() @ X[:, :d].float()
    c2 = torch.zeros(d, H * hd, device="cuda")
    ops.gemm_tn(X, A2, c2, H * hd, 1, rows, d, [(0, 0, 0, H * hp)], (0, 0), (hd, hp))
    print(f"two_stage={mode} head-padded rows: rel err {((c1 - r1).norm() / r1.norm()).item():.3e}; "
          f"columns: {((c2 - r1.t()).norm() / r1.norm()).item():.3e}", flush=True)

for (M, Na, Dp, nlen, k) in [(39424, 270, 272, 270, 1), (39424, 270, 272, 270, 3), (39424, 270, 272, 270, 5), (39424, 960, 272, 270, 1),
                             (39424, 270, 320, 320, 1), (39424, 128, 272, 270, 8), (39424, 128, 272, 270, 16), (70144, 150, 160, 150, 1),
                             (70144, 150, 160, 150, 2), (70144, 150, 160, 150, 3), (70144, 480, 160, 150, 1)]:
    A, full, X, Cm, segs, pl = make(M, Na, Dp, nlen, k)
    fl = 2.0 * M * Na * nlen * k
    res = []
    two_stage(False)
    for dbg in (0, 1, 2):
        lib.csi_set_tn_debug(dbg)
        ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
        res.append(f"atomics dbg{dbg} {ms*1e3:6.1f} us")
    lib.csi_set_tn_debug(0)
    two_stage(True)
    ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
    res.append(f"two-stage {ms*1e3:6.1f} us {fl/ms/1e9:6.1f} TF/s")
    lib.csi_set_tn_debug(32)
    ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
    res.append(f"(stage one only {ms*1e3:6.1f} us)")
    lib.csi_set_tn_debug(64)
    ms = timeit(lambda: ops.gemm_tn(A, X, Cm, nlen * k, k, M, Na, segs))
    res.append(f"(stage two only {ms*1e3:6.1f} us)")
    lib.csi_set_tn_debug(0)
    print(f"M={M} Na={Na} nlen={nlen} k={k}: " + " | ".join(res), flush=True)
