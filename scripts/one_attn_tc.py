"""A few launches of one attention kernel variant at B=256 (for ncu): argv = kind(tc|tc2|mma|tcbwd|mmabwd) L d [reps]."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps, _p, _ld
kind, L, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
B, H, HALO, GUARD = 256, 10, 2, 16
ops = NativeOps(torch.device("cuda", 0)); lib = ops.lib
hd = d // H; hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
Lp = L + 2 * HALO
def mk(nsec, scale):
    full = torch.zeros(B * Lp + 2 * GUARD, nsec * H * hp, dtype=torch.bfloat16, device="cuda")
    body = full[GUARD:GUARD + B * Lp]
    body.view(B, Lp, nsec * H, hp)[:, HALO:HALO + L, :, :hd] = (torch.randn(B, L, nsec * H, hd, device="cuda") * scale).to(torch.bfloat16)
    return body
qkv, do, o, dqkv = mk(3, 0.5), mk(1, 0.5), mk(1, 0.0), mk(3, 0.0)
lse = torch.zeros(B * H * L, device="cuda"); dbias = torch.zeros(3 * d, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fwd = {"tc": lib.csi_attn_fwd_tc, "mma": lib.csi_attn_fwd_mma}
lib.csi_attn_fwd_mma(_p(qkv), _ld(qkv), _p(o), _ld(o), _p(lse), B, L, d, H, hp, HALO, st)
for _ in range(reps):
    if kind in fwd:
        rc = fwd[kind](_p(qkv), _ld(qkv), _p(o), _ld(o), _p(lse), B, L, d, H, hp, HALO, st)
    else:
        fn = lib.csi_attn_bwd_tc if kind == "tcbwd" else lib.csi_attn_bwd_mma
        rc = fn(_p(qkv), _ld(qkv), _p(o), _ld(o), _p(do), _ld(do), _p(dqkv), _ld(dqkv), _p(lse), B, L, d, H, hp, HALO, _p(dbias), st)
    assert rc == 0, lib.csi_last_error()
torch.cuda.synchronize()
print("ok")
