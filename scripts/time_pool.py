"""Device time of csi_pool_dual (augment on / off) at B=256, T=3000, F=270|540, ten launches replayed from a CUDA graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
ops = NativeOps(torch.device("cuda", 0))
B, T, HALO, GUARD = 256, 3000, 2, 16
for F in (270, 540):
    L = T // 20
    x = torch.rand(B, T, F, device="cuda") * 20
    Dl, Dr = (F + 15) // 16 * 16, (L + 15) // 16 * 16
    left = torch.zeros(B * (L + 2 * HALO) + 2 * GUARD, Dl, device="cuda")[GUARD:-GUARD]
    right = torch.zeros(B * (F + 2 * HALO) + 2 * GUARD, Dr, device="cuda")[GUARD:-GUARD]
    pe = torch.randn(L, Dl, device="cuda")
    rng = torch.tensor([5, 3], dtype=torch.int64, device="cuda")
    for aug in (0, 1):
        fn = lambda: ops.pool_dual(x, None, None, B, T, F, pe, left, right, HALO, aug, rng)
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        nb = B * T * F * 4 + 2 * B * L * F * 4
        print(f"F={F} augment={aug}: {us:.1f} us  {nb / us / 1e3:.0f} GB/s", flush=True)
