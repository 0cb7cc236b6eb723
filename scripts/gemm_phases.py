"""Phase clocks of one mid-grid CTA of the NT GEMM (csi_set_gemm_debug): argv = M N Dp k mode(single|pair|default) [res]."""
import sys, os, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_csi_b200.ops import NativeOps
M, N, Dp, k = [int(a) for a in sys.argv[1:5]]
mode = sys.argv[5] if len(sys.argv) > 5 else "default"
RES = len(sys.argv) > 6 and sys.argv[6] == "res"
ops = NativeOps(torch.device("cuda", 0)); lib = ops.lib
lib.csi_set_gemm_pair({"single": 0, "pair": 2, "default": 1}[mode])
lib.csi_set_gemm_resident(1 if mode == "default" else 0)
GUARD = 16
full = torch.randn(M + 2 * GUARD, Dp, device="cuda").to(torch.bfloat16)
A = full[GUARD:GUARD + M]
W = (torch.randn(N, k * Dp, device="cuda") / math.sqrt(k * Dp)).to(torch.bfloat16)
Cm = torch.zeros(M, (N + 15) // 16 * 16, dtype=torch.float32 if RES else torch.bfloat16, device="cuda")
bias = torch.randn(N, device="cuda") if RES else None
res = torch.randn_like(Cm) if RES else None
rng = torch.tensor([1234, 7], dtype=torch.int64, device="cuda") if RES else None
pl = (k - 1) // 2
segs = [(j - pl, 0, j * Dp, Dp) for j in range(k)]
def go():
    ops.gemm_nt(A, W, Cm, M, N, segs, bias, res, 0.1 if RES else 0.0, 3, rng)
for _ in range(3):
    go()
dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.csi_set_gemm_debug(C.c_void_p(dbg.data_ptr()))
go()
torch.cuda.synchronize()
lib.csi_set_gemm_debug(C.c_void_p(0))
t = dbg.cpu().tolist()
t0 = min(v for v in t if v)
print(f"M={M} N={N} K={k}x{Dp} {mode} res={RES}:  MMA [tile top | acc free | first A | committed]   EPI warp0 [top | acc full | ld done | buf free | (sts done | fenced) | stored | end]   PROD [tile start]")
for i in range(12):
    m = [v - t0 if v else -1 for v in t[8 * i: 8 * i + 4]]
    e = [v - t0 if v else -1 for v in t[1000 + 8 * i: 1000 + 8 * i + 8]]
    pr = t[2000 + 8 * i] - t0 if t[2000 + 8 * i] else -1
    if m[0] < 0 and e[0] < 0:
        break
    print(f"  tile {i:2d}  mma {m[0]:6d} {m[1]:6d} {m[2]:6d} {m[3]:6d}   epi {e[0]:6d} {e[1]:6d} {e[2]:6d} {e[3]:6d} ({e[6]:6d} {e[7]:6d}) {e[4]:6d} {e[5]:6d}   prod {pr:6d}")
