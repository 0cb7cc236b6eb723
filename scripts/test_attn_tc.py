"""tcgen05 attention (attention_tc.cu) against torch SDPA autograd and the earlier mma.sync kernel, plus CUDA-event timing.

    python scripts/test_attn_tc.py [B_time]

Every configuration is checked on its own (errors are printed, not raised) so that one GPU call reports all of them."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_modal_csi_b200.ops import NativeOps, _p, _ld

BT = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops = NativeOps(torch.device("cuda", 0))
lib = ops.lib
HALO, GUARD = 2, 16


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mk(B, L, H, hd, hp, nsec, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    Lp = L + 2 * HALO
    rows = B * Lp
    full = torch.zeros(rows + 2 * GUARD, nsec * H * hp, dtype=torch.bfloat16, device="cuda")
    body = full[GUARD:GUARD + rows]
    v = torch.randn(B, L, nsec * H, hd, device="cuda", generator=g) * scale
    body.view(B, Lp, nsec * H, hp)[:, HALO:HALO + L, :, :hd] = v.to(torch.bfloat16)
    return full, body


def valid(body, B, L, nh, hp, hd):
    return body.view(B, L + 2 * HALO, nh, hp)[:, HALO:HALO + L, :, :hd].float()


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def run_fwd(kind, qkv, o, lse, B, L, d, H, hp):
    fn = lib.csi_attn_fwd_tc2 if kind == "tc2" else {"tc": lib.csi_attn_fwd_tc, "mma": lib.csi_attn_fwd_mma}[kind]
    rc = fn(_p(qkv), _ld(qkv), _p(o), _ld(o), _p(lse), B, L, d, H, hp, HALO, st())
    if rc:
        raise RuntimeError(lib.csi_last_error().decode())


def run_bwd(kind, qkv, o, do, dqkv, lse, B, L, d, H, hp, dbias):
    fn = lib.csi_attn_bwd_tc if kind == "tc" else lib.csi_attn_bwd_mma
    rc = fn(_p(qkv), _ld(qkv), _p(o), _ld(o), _p(do), _ld(do), _p(dqkv), _ld(dqkv), _p(lse), B, L, d, H, hp, HALO, _p(dbias), st())
    if rc:
        raise RuntimeError(lib.csi_last_error().decode())


def check(B, L, d, H=10, bwd=True):
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    if not lib.csi_attn_tc_ok(L, d, H, hp):
        print(f"[skip] L={L} d={d}: not eligible for the tcgen05 kernel", flush=True)
        return
    _, qkv = mk(B, L, H, hd, hp, 3, 1)
    _, do = mk(B, L, H, hd, hp, 1, 2)
    t = valid(qkv, B, L, 3 * H, hp, hd).reshape(B, L, 3, H, hd)
    q, k, v = [t[:, :, w].transpose(1, 2).clone().requires_grad_(True) for w in range(3)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    gdo = valid(do, B, L, H, hp, hd).transpose(1, 2)
    ref.backward(gdo)
    ref_o = ref.transpose(1, 2)
    ref_lse = torch.logsumexp((q @ k.transpose(-1, -2)) / hd ** 0.5, dim=-1)       # [B,H,L]
    out = {}
    for kind in ("mma", "tc") + (("tc2",) if hasattr(lib, "csi_attn_fwd_tc2") else ()):
        if kind == "tc2" and not lib.csi_attn_tc2_ok(L, d, H, hp):
            continue
        fo, o = mk(B, L, H, hd, hp, 1, 0, 0.0)
        lse = torch.zeros(B * H * L, device="cuda")
        try:
            run_fwd(kind, qkv, o, lse, B, L, d, H, hp)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"[FAIL] fwd {kind} L={L} d={d}: {e}", flush=True)
            continue
        eo = rel(valid(o, B, L, H, hp, hd), ref_o)
        el = rel(lse.view(B, H, L), ref_lse)
        pad = float(o.view(B, L + 2 * HALO, H, hp)[:, :, :, hd:].float().abs().max()) if hp > hd else 0.0
        halo = float(o.view(B, L + 2 * HALO, -1)[:, :HALO].float().abs().max())
        guard = float(fo[:GUARD].float().abs().max() + fo[-GUARD:].float().abs().max())
        msg = f"fwd {kind:3s} B={B} L={L} d={d} hp={hp}: o {eo:.2e} lse {el:.2e} pad {pad} halo {halo} guard {guard}"
        if bwd and kind != "tc2" and (kind == "mma" or hasattr(lib, "csi_attn_bwd_tc")):
            _, dqkv = mk(B, L, H, hd, hp, 3, 0, 0.0)
            dbias = torch.full((3 * d,), 0.5, device="cuda")
            try:
                if kind == "tc" and not lib.csi_attn_bwd_tc_ok(L, d, H, hp):
                    raise RuntimeError("bwd not eligible")
                run_bwd(kind, qkv, o, do, dqkv, lse, B, L, d, H, hp, dbias)
                torch.cuda.synchronize()
                g = valid(dqkv, B, L, 3 * H, hp, hd).reshape(B, L, 3, H, hd)
                errs = [rel(g[:, :, w].transpose(1, 2), x.grad) for w, x in enumerate((q, k, v))]
                cs = g.reshape(B * L, 3, H, hd).sum(0).reshape(-1) + 0.5
                padg = float(dqkv.view(B, L + 2 * HALO, 3 * H, hp)[:, :, :, hd:].float().abs().max()) if hp > hd else 0.0
                msg += f" | bwd dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e} dbias {rel(dbias, cs):.2e} pad {padg}"
            except Exception as e:
                msg += f" | bwd FAIL {e}"
        print(msg, flush=True)


def timeit(B, L, d, H=10):
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    _, qkv = mk(B, L, H, hd, hp, 3, 1, 0.5)
    _, do = mk(B, L, H, hd, hp, 1, 2, 0.5)
    _, o = mk(B, L, H, hd, hp, 1, 0, 0.0)
    _, dqkv = mk(B, L, H, hd, hp, 3, 0, 0.0)
    lse = torch.zeros(B * H * L, device="cuda")
    dbias = torch.zeros(3 * d, device="cuda")
    line = f"time B={B} L={L} d={d} hp={hp}:"
    for kind in ("mma", "tc") + (("tc2",) if hasattr(lib, "csi_attn_fwd_tc2") else ()):
        for name, fn in (("fwd", lambda: run_fwd(kind, qkv, o, lse, B, L, d, H, hp)),
                         ("bwd", lambda: run_bwd(kind, qkv, o, do, dqkv, lse, B, L, d, H, hp, dbias))):
            try:
                if kind == "tc2" and (name == "bwd" or not lib.csi_attn_tc2_ok(L, d, H, hp)):
                    continue
                if kind == "tc" and (not lib.csi_attn_tc_ok(L, d, H, hp) or (name == "bwd" and not lib.csi_attn_bwd_tc_ok(L, d, H, hp))):
                    continue
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                line += f"  {kind}.{name} {e0.elapsed_time(e1) / 20 * 1e3:.1f} us"
            except Exception as e:
                line += f"  {kind}.{name} FAIL({e})"
    print(line, flush=True)


def phases(B, L, d, H=10):
    """Per-phase clocks of one mid-grid CTA (last warp): where a (head, tile) job spends its time."""
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    _, qkv = mk(B, L, H, hd, hp, 3, 1, 0.5)
    _, o = mk(B, L, H, hd, hp, 1, 0, 0.0)
    lse = torch.zeros(B * H * L, device="cuda")
    dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
    run_fwd("tc", qkv, o, lse, B, L, d, H, hp)
    lib.csi_set_attn_debug(C.c_void_p(dbg.data_ptr()))
    run_fwd("tc", qkv, o, lse, B, L, d, H, hp)
    torch.cuda.synchronize()
    lib.csi_set_attn_debug(C.c_void_p(0))
    t = dbg.cpu().tolist()
    print(f"phases L={L} d={d}:", flush=True)
    job = 0
    names = ["S+wait", "pass1", "pass2", "bar+PV", "epi"]
    while 16 + job * 8 + 5 < 512 and t[16 + job * 8 + 5] and job < 12:
        v = t[16 + job * 8: 22 + job * 8]
        prev = t[16 + job * 8 - 3] if job else t[0]
        print(f"  job {job}: gap {v[0] - prev} " + " ".join(f"{n} {v[k + 1] - v[k]}" for k, n in enumerate(names)), flush=True)
        job += 1


def phases_bwd(B, L, d, H=10):
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    _, qkv = mk(B, L, H, hd, hp, 3, 1, 0.5)
    _, do = mk(B, L, H, hd, hp, 1, 2, 0.5)
    _, o = mk(B, L, H, hd, hp, 1, 0, 0.0)
    _, dqkv = mk(B, L, H, hd, hp, 3, 0, 0.0)
    lse = torch.zeros(B * H * L, device="cuda")
    dbias = torch.zeros(3 * d, device="cuda")
    run_fwd("tc", qkv, o, lse, B, L, d, H, hp)
    dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
    run_bwd("tc", qkv, o, do, dqkv, lse, B, L, d, H, hp, dbias)
    lib.csi_set_attn_bwd_debug(C.c_void_p(dbg.data_ptr()))
    run_bwd("tc", qkv, o, do, dqkv, lse, B, L, d, H, hp, dbias)
    torch.cuda.synchronize()
    lib.csi_set_attn_bwd_debug(C.c_void_p(0))
    t = dbg.cpu().tolist()
    print(f"bwd phases L={L} d={d}:  MMA warp per job: [issue S/dP] [wait drained] [wait P/dS] [issue grads]; loop gap", flush=True)
    for j in range(10):
        v = t[8 * j: 8 * j + 4]
        if not v[3]:
            break
        nxt = t[8 * (j + 1)]
        print(f"  mma job {j}: issueS {v[1] - v[0]} waitPdS {v[2] - v[1]} issueG {v[3] - v[2]} gap {nxt - v[3] if nxt else 0}", flush=True)
    print("  worker (warp 0) per key tile: [wait S/dP] [compute+arrive] [wait grads] [drain]", flush=True)
    for j in range(10):
        v = t[500 + 8 * j: 500 + 8 * j + 5]
        if not v[4]:
            break
        print(f"  wrk kt {j}: waitS {v[1] - v[0]} compute {v[2] - v[1]} waitG {v[3] - v[2]} drain {v[4] - v[3]}", flush=True)


def phases2(B, L, d, H=10):
    """attention_tc2.cu: clocks of warpgroup 0 / 1 (quarter 0) and of the MMA warp of one mid-grid CTA."""
    hd = d // H
    hp = 16 if hd <= 16 else 32 if hd <= 32 else 64
    _, qkv = mk(B, L, H, hd, hp, 3, 1, 0.5)
    _, o = mk(B, L, H, hd, hp, 1, 0, 0.0)
    lse = torch.zeros(B * H * L, device="cuda")
    dbg = torch.zeros(2048, dtype=torch.int64, device="cuda")
    run_fwd("tc2", qkv, o, lse, B, L, d, H, hp)
    lib.csi_set_attn2_debug(C.c_void_p(dbg.data_ptr()))
    run_fwd("tc2", qkv, o, lse, B, L, d, H, hp)
    torch.cuda.synchronize()
    lib.csi_set_attn2_debug(C.c_void_p(0))
    t = dbg.cpu().tolist()
    t0 = min(v for v in t if v)
    print(f"phases2 L={L} d={d}: job: WG [wait S | softmax | wait O | drain]   MMA [S issued at | P seen | PV issued]", flush=True)
    for n in range(14):
        w = t[8 * n: 8 * n + 5]
        m = t[1000 + 8 * n: 1000 + 8 * n + 6]
        if not w[4]:
            break
        print(f"  job {n:2d} wg{n & 1}: start {w[0] - t0:6d} waitS {w[1] - w[0]:5d} softmax {w[2] - w[1]:5d} waitO {w[3] - w[2]:5d} drain {w[4] - w[3]:5d}"
              f" | mma: loop {m[1] - t0:6d} S-issued {m[2] - t0:6d} P-seen {m[3] - t0 if m[3] else 0:6d} PV-issued {m[5] - t0 if m[5] else 0:6d}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "phases2":
        for (L, d) in [(150, 270), (150, 540)]:
            phases2(BT, L, d)
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "phases_bwd":
        for (L, d) in [(150, 270), (270, 150), (150, 540)]:
            phases_bwd(BT, L, d)
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "phases":
        for (L, d) in [(150, 270), (270, 150), (150, 540)]:
            phases(BT, L, d)
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2].startswith("ts"):
        for mode in (1, 2, 3):
            lib.csi_set_attn_teams(mode)
            print(f"--- teams <= {mode}", flush=True)
            for (B, L, d) in [(2, 20, 30), (3, 150, 270), (2, 270, 150), (2, 150, 540)]:
                check(B, L, d, bwd=False)
            for (L, d) in [(150, 270), (270, 150)]:
                timeit(BT, L, d)
        sys.exit(0)
    for (B, L, d) in [(2, 20, 30), (2, 30, 20), (3, 150, 270), (2, 270, 150), (2, 150, 540), (1, 540, 150), (2, 128, 320), (2, 129, 160), (1, 16, 640)]:
        check(B, L, d)
    for (L, d) in [(150, 270), (270, 150), (150, 540), (540, 150)]:
        timeit(BT, L, d)
