"""Times csi_layernorm_bwd / fwd on the THAT shapes (device time per call, GB/s of algorithmic bytes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
ops = NativeOps(torch.device("cuda", 0))
HALO, GUARD = 2, 16
def buf(B, L, ld, dt):
    full = torch.randn((B * (L + 2 * HALO)) + 2 * GUARD, ld, device="cuda").to(dt)
    return full[GUARD:GUARD + B * (L + 2 * HALO)]
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (B, L, d) in [(256, 150, 270), (256, 270, 150), (256, 150, 540), (256, 540, 150)]:
    Dp = (d + 15) // 16 * 16
    rows = B * (L + 2 * HALO)
    dy = buf(B, L, Dp, torch.bfloat16); x = buf(B, L, Dp, torch.float32); dres = buf(B, L, Dp, torch.float32)
    dx = buf(B, L, Dp, torch.float32); dxm = buf(B, L, Dp, torch.bfloat16); y = buf(B, L, Dp, torch.bfloat16)
    gam = torch.randn(d, device="cuda"); bet = torch.randn(d, device="cuda")
    mean = torch.zeros(rows, device="cuda"); rstd = torch.ones(rows, device="cuda")
    dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda")
    rng = torch.tensor([1, 2], dtype=torch.int64, device="cuda")
    t_f = timeit(lambda: ops.layernorm_fwd(x, gam, bet, y, mean, rstd, B, L, d, HALO, 1e-6))
    t_b = timeit(lambda: ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dx, dxm, 0.1, 7, rng, dg, db, B, L, d, HALO))
    t_b0 = timeit(lambda: ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dx, None, 0.0, 0, rng, dg, db, B, L, d, HALO))
    n = B * L * d
    print(f"B={B} L={L} d={d}: fwd {t_f*1e3:6.1f} us {n*6/t_f/1e6:6.0f} GB/s | bwd+mask {t_b*1e3:6.1f} us {n*16/t_b/1e6:6.0f} GB/s | bwd {t_b0*1e3:6.1f} us {n*14/t_b0/1e6:6.0f} GB/s", flush=True)
