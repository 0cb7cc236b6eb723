"""Times attn_fwd / attn_bwd (CUDA events) on the THAT stream shapes: argv = [B]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops = NativeOps(torch.device("cuda", 0))
HALO, GUARD = 2, 16
for (L, d, H) in [(150, 270, 10), (270, 150, 10), (150, 540, 10), (540, 150, 10)]:
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    rows = B * (L + 2 * HALO)
    def buf(ld, fill=True):
        full = (torch.randn(rows + 2 * GUARD, ld, device="cuda") * (0.5 if fill else 0)).to(torch.bfloat16)
        return full[GUARD:GUARD + rows]
    qkv, o, do, dqkv = buf(3 * H * hp), buf(H * hp, False), buf(H * hp), buf(3 * H * hp, False)
    lse = torch.zeros(B * H * L, device="cuda")
    dbias = torch.zeros(3 * d, device="cuda")
    res = []
    for fn in (lambda: ops.attn_fwd(qkv, o, lse, B, L, d, H, hp, HALO),
               lambda: ops.attn_bwd(qkv, o, do, dqkv, lse, B, L, d, H, hp, HALO, dbias)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    fl = 4.0 * B * H * L * L * hd
    print(f"L={L} d={d} hp={hp}: fwd {res[0]:.1f} us ({fl / res[0] / 1e6:.0f} TF/s)  bwd {res[1]:.1f} us ({2.5 * fl / res[1] / 1e6:.0f} TF/s)", flush=True)
