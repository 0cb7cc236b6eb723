"""Times attn_fwd / attn_bwd on the THAT stream shapes, ten launches replayed from a CUDA graph: argv = [B]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops = NativeOps(torch.device("cuda", 0))
HALO, GUARD = 2, 16


def timeit(fn, reps=10):
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
    return e0.elapsed_time(e1) / reps * 1e3


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
    f = timeit(lambda: ops.attn_fwd(qkv, o, lse, B, L, d, H, hp, HALO))
    b = timeit(lambda: ops.attn_bwd(qkv, o, do, dqkv, lse, B, L, d, H, hp, HALO, dbias))
    fl = 4.0 * B * H * L * L * hd
    print(f"L={L} d={d} hp={hp}: fwd {f:.1f} us ({fl / f / 1e6:.0f} TF/s)  bwd {b:.1f} us ({2.5 * fl / b / 1e6:.0f} TF/s)", flush=True)
