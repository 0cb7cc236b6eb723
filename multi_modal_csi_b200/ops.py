"""ctypes binding of ``libcsi_that.so`` (include/csi_that.h).

This is the whole Python side of the C ABI: every method takes torch CUDA tensors, passes raw device
pointers, sizes and the *current* CUDA stream, and checks the returned ``csi_status``.  There is no
fallback: a missing library, a non-CUDA device or a non-sm_100 GPU raises at construction.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsi_that.so")
ABI_VERSION = 5


class Seg(C.Structure):
    _fields_ = [("a_row_shift", C.c_int), ("a_col_off", C.c_int), ("b_col_off", C.c_int), ("klen", C.c_int)]


class Band(C.Structure):
    """csi_band: output columns [n_lo, n_hi) a segment of csi_gemm_nt_banded contributes to."""
    _fields_ = [("n_lo", C.c_int), ("n_hi", C.c_int)]


class SegTN(C.Structure):
    _fields_ = [("b_row_shift", C.c_int), ("b_col_off", C.c_int), ("c_off", C.c_int), ("nlen", C.c_int)]


class Ptr3(C.Structure):
    _fields_ = [("p", C.c_void_p * 3)]


class Grp(C.Structure):
    """csi_grp: head padding (valid, pad); (0, 0) = none."""
    _fields_ = [("valid", C.c_int), ("pad", C.c_int)]


NO_GRP = (0, 0)


class PackEntry(C.Structure):
    _fields_ = [("src_off", C.c_longlong), ("dst_off", C.c_longlong), ("N", C.c_int), ("C", C.c_int),
                ("k", C.c_int), ("ld", C.c_int), ("mode", C.c_int), ("P", C.c_int), ("seg_base", C.c_int),
                ("gn", Grp), ("gc", Grp), ("reserved", C.c_int)]


_lib = None


def load_library():
    """Loads libcsi_that.so (built by ``__graft_entry__.build()`` / ``make -C multi_modal_csi_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    lib.csi_last_error.restype = C.c_char_p
    if lib.csi_abi_version() != ABI_VERSION:
        raise RuntimeError("libcsi_that.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


EXPORTS = [
    "csi_last_error", "csi_abi_version", "csi_device_arch", "csi_pool_dual", "csi_gauss_pe_fwd", "csi_gauss_pe_bwd",
    "csi_layernorm_fwd", "csi_layernorm_bwd", "csi_gemm_nt", "csi_gemm_tn", "csi_colsum_tokens", "csi_attn_fwd",
    "csi_attn_bwd", "csi_bn_stats", "csi_bn_finalize", "csi_bn_eval_prepare", "csi_bn_act_fwd",
    "csi_bn_act_bwd_reduce", "csi_bn_act_bwd_dz", "csi_head_reduce_fwd", "csi_head_reduce_bwd", "csi_dropout_rows",
    "csi_bce_logits", "csi_smooth_l1", "csi_perm_ce", "csi_predict_counts", "csi_adam_flat", "csi_advance_counters", "csi_pack_weights", "csi_fill_f32",
    "csi_set_force_simt", "csi_set_strict_tc", "csi_dispatch_counts", "csi_set_attn_impl",
    "csi_fill_f64", "csi_copy_f32", "csi_nhwc_stats", "csi_bn2d_finalize", "csi_im2col_bn", "csi_col2im", "csi_act_drop_fwd",
    "csi_bn2d_bwd_reduce", "csi_bn2d_bwd_apply", "csi_pool_bn_fwd", "csi_conv2d_pack", "csi_conv2d_unpack_grad", "csi_bn0_grads",
    "csi_gather_aug", "csi_gemm_tn_workspace", "csi_gemm_nt_banded",
]

# Workspaces of the two-stage weight-gradient reduction (csi_gemm_tn_workspace), one per (device, stream) that ever issued a
# weight gradient; they live as long as the process, so the library never holds a pointer to freed memory.
# Largest need: one wave of CTAs x 128 rows x 512 TMEM columns of fp32 = 38.8 MB on a 148-SM part.
TN_TWO_STAGE = os.environ.get("CSI_TN_TWO_STAGE", "1") != "0"
_TN_WS = {}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _ld(t):
    return 0 if t is None else int(t.stride(0))


class NativeOps:
    """Thin, stateless wrapper: one method per C entry point."""

    def __init__(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("libcsi_that.so kernels need a CUDA device (B200, sm_100a); got " + str(device))
        self.lib = load_library()
        self.device = device
        arch = self.lib.csi_device_arch(C.c_int(device.index or 0))
        if arch < 0:
            raise RuntimeError("csi_device_arch failed: " + self.lib.csi_last_error().decode())
        self.arch = arch
        self.launches = 0
        self._prof = None
        self._only = None
        self._tag = ""
        self.last_calls = []

    # -------------------------------------------------------------- helpers
    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc, name):
        self.launches += 1
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.lib.csi_last_error().decode()}")

    @staticmethod
    def _ptr3(ts):
        p = Ptr3()
        for i in range(3):
            p.p[i] = ts[i].data_ptr() if i < len(ts) else 0
        return p


    # -------------------------------------------------------------- launch + optional per-op device timing
    def _call(self, name, *args, flops=0, nbytes=0, launches=1):
        short = name[4:]
        if short == "gemm_nt_banded":
            short = "gemm_nt"                         # one op family in the profile tables
        prof = self._prof is not None and (self._only is None or short in self._only)
        if prof:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
        rc = getattr(self.lib, name)(*args, self._st())
        self.launches += launches
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.lib.csi_last_error().decode()}")
        if prof:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.device))
            self._prof.append((short, e0, e1, flops, nbytes, self._tag))
        self._tag = ""

    def start_profile(self, only=None):
        """Record a CUDA-event pair around every launch of the named ops (all ops when only is None)."""
        self._prof, self._only = [], (set(only) if only else None)

    def stop_profile(self):
        """-> {op: (total_ms, launches, {"flops": f, "bytes": b})}; synchronises."""
        torch.cuda.synchronize(self.device)
        out = {}
        self.last_calls = [(name, e0.elapsed_time(e1), fl, nb, tag) for name, e0, e1, fl, nb, tag in (self._prof or [])]
        for name, e0, e1, fl, nb, _tag in (self._prof or []):
            ms, n, w = out.get(name, (0.0, 0, {"flops": 0, "bytes": 0}))
            w["flops"] += fl
            w["bytes"] += nb
            out[name] = (ms + e0.elapsed_time(e1), n + 1, w)
        self._prof = None
        return out

    def set_force_simt(self, on: bool):
        self.lib.csi_set_force_simt(C.c_int(1 if on else 0))

    def set_strict_tc(self, on: bool):
        """Strict mode: a bf16 contraction whose shape the tcgen05 kernel cannot take is an error instead of an FFMA run."""
        self.lib.csi_set_strict_tc(C.c_int(1 if on else 0))

    def dispatch_counts(self, reset: bool = False):
        """{"tcgen05": n, "ffma_fallback": n, "mma_sync": n}: bf16 contraction / attention calls by the kernel class that served them."""
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self.lib.csi_dispatch_counts(C.byref(a), C.byref(b), C.byref(c), C.c_int(1 if reset else 0))
        return {"tcgen05": a.value, "ffma_fallback": b.value, "mma_sync": c.value}

    def set_attn_impl(self, fwd_mode: int = 0, bwd_mode: int = 0):
        """0 = measured-best per shape, 1 = tcgen05/TMEM attention wherever eligible, 2 = mma.sync only."""
        self.lib.csi_set_attn_impl(C.c_int(fwd_mode), C.c_int(bwd_mode))

    def make_pack_table(self, entries, device):
        arr = (PackEntry * len(entries))()
        for i, e in enumerate(entries):
            arr[i] = PackEntry(*e[:9], Grp(*e[9]), Grp(*e[10]), 0)
        raw = bytes(arr)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(device)
        return t


    # -------------------------------------------------------------- algorithmic work per launch (roofline numerators)
    alg_scale = 1.0     # set by the engine: (valid tokens / padded rows) * (true channels / padded channels)
    alg_flops = None    # set by the engine before a contraction call: its algorithmic FLOPs (true dims, no padding)

    def _work(self, name, a):
        """Algorithmic FLOPs / minimum HBM bytes of one launch (DESIGN.md lists the formulas)."""
        es = lambda t: 0 if t is None else t.element_size()
        if name in ("gemm_nt", "gemm_tn") and self.alg_flops is not None:
            fl, self.alg_flops = self.alg_flops, None
        else:
            fl = None
        if name == "gemm_nt":
            k = sum(s[3] for s in a["segs"])
            return {"flops": fl if fl is not None else int(2 * a["M"] * a["N"] * k * self.alg_scale),
                    "nbytes": a["M"] * k // max(1, len(a["segs"])) * es(a["A"]) + a["N"] * k * es(a["Bw"]) + a["M"] * a["N"] * es(a["Cm"])}
        if name == "gemm_tn":
            n = sum(s[3] for s in a["segs"])
            return {"flops": fl if fl is not None else int(2 * a["M"] * a["Na"] * n * (self.alg_scale if n > 16 else 1.0)), "nbytes": a["M"] * (a["Na"] + n // max(1, len(a["segs"]))) * es(a["A"])}
        if name == "attn_fwd":
            hd = a["d"] // a["H"]
            return {"flops": 4 * a["B"] * a["H"] * a["L"] * a["L"] * hd, "nbytes": a["B"] * a["L"] * a["d"] * 4 * es(a["qkv"])}
        if name == "attn_bwd":
            hd = a["d"] // a["H"]
            return {"flops": 10 * a["B"] * a["H"] * a["L"] * a["L"] * hd, "nbytes": a["B"] * a["L"] * a["d"] * 8 * es(a["qkv"])}
        if name == "pool_dual":
            return {"nbytes": a["B"] * a["T"] * a["F"] * 4 + 2 * a["B"] * (a["T"] // 20) * a["F"] * 4}
        if name == "layernorm_fwd":
            return {"nbytes": a["B"] * a["L"] * a["d"] * (4 + es(a["y"]))}
        if name == "layernorm_bwd":
            return {"nbytes": a["B"] * a["L"] * a["d"] * (es(a["dy"]) + 4 + 4 + (4 if a["dres"] is not None else 0) + es(a["dxm"]))}
        if name in ("bn_stats", "colsum_tokens"):
            t = a.get("z", a.get("A"))
            return {"nbytes": a["B"] * a["L"] * a["ncols"] * es(t)}
        if name == "bn_act_fwd":
            return {"nbytes": a["B"] * a["L"] * a["d"] * (a["nbr"] * es(a["z"]) + 8)}
        if name == "bn_act_bwd_reduce":
            return {"nbytes": a["B"] * a["L"] * a["d"] * (a["nbr"] * es(a["z"]) + 4)}
        if name == "bn_act_bwd_dz":
            return {"nbytes": a["B"] * a["L"] * a["d"] * (2 * a["nbr"] * es(a["z"]) + 4)}
        if name == "adam_flat":
            return {"nbytes": 28 * a["n"]}
        if name == "gauss_pe_bwd":
            return {"nbytes": a["B"] * a["L"] * a["F"] * 4, "launches": 3}
        if name in ("head_reduce_fwd", "head_reduce_bwd"):
            return {"nbytes": a["B"] * a["L"] * a["N"] * es(a["p"]) * (2 if name.endswith("bwd") else 1)}
        return {}

    # -------------------------------------------------------------- ops
    def pack_weights(self, params, packed, table, n_entries, max_elems):
        wk = self._work("pack_weights", locals())
        self._call("csi_pack_weights", _p(params), _p(packed), _dt(packed), _p(table), n_entries, max_elems, **wk)

    def pool_dual(self, x, offs, lens, B, T, F, pe, left, right, halo, augment, rng):
        wk = self._work("pool_dual", locals())
        self._call("csi_pool_dual", _p(x), _p(offs), _p(lens), B, T, F, _p(pe), _ld(pe), _p(left), _ld(left),
                                        _p(right), _ld(right), halo, int(augment), _p(rng), **wk)

    def gauss_pe_fwd(self, pos, mu, sigma, emb, L, K, F, w, pe):
        wk = self._work("gauss_pe_fwd", locals())
        self._call("csi_gauss_pe_fwd", _p(pos), _p(mu), _p(sigma), _p(emb), L, K, F, _p(w), _p(pe), _ld(pe), **wk)

    def gauss_pe_bwd(self, dleft, B, halo, w, pos, mu, sigma, emb, L, K, F, dpe_ws, demb, dmu, dsigma):
        wk = self._work("gauss_pe_bwd", locals())
        self._call("csi_gauss_pe_bwd", _p(dleft), _ld(dleft), B, halo, _p(w), _p(pos), _p(mu), _p(sigma), _p(emb),
                                           L, K, F, _p(dpe_ws), _ld(dpe_ws), _p(demb), _p(dmu), _p(dsigma), **wk)

    def layernorm_fwd(self, x, gamma, beta, y, mean, rstd, B, L, d, halo, eps):
        wk = self._work("layernorm_fwd", locals())
        self._call("csi_layernorm_fwd", _p(x), _ld(x), _p(gamma), _p(beta), _p(y), _ld(y), _dt(y), _p(mean),
                                            _p(rstd), B, L, d, halo, C.c_float(eps), **wk)

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, dres, dx, dxm, drop_p, drop_site, rng, dgamma, dbeta,
                      B, L, d, halo):
        wk = self._work("layernorm_bwd", locals())
        self._call("csi_layernorm_bwd", _p(dy), _ld(dy), _dt(dy), _p(x), _ld(x), _p(gamma), _p(mean), _p(rstd),
                                            _p(dres), _ld(dres), _p(dx), _ld(dx), _p(dxm), _ld(dxm),
                                            0 if dxm is None else _dt(dxm), C.c_float(drop_p), C.c_uint(drop_site),
                                            _p(rng), _p(dgamma), _p(dbeta), B, L, d, halo, **wk)

    def gemm_nt(self, A, Bw, Cm, M, N, segs, bias, residual, drop_p, drop_site, rng):
        wk = self._work("gemm_nt", locals())
        if self._prof is not None:
            self._tag = f"M={M} N={N} K={sum(s[3] for s in segs)} nseg={len(segs)} out={Cm.dtype} res={residual is not None} drop={drop_p}"
        arr = (Seg * len(segs))(*[Seg(*s) for s in segs])
        self._call("csi_gemm_nt", _p(A), _ld(A), _p(Bw), _ld(Bw), _dt(A), _p(Cm), _ld(Cm), _dt(Cm), M, N, arr,
                                      len(segs), _p(bias), _p(residual), _ld(residual), C.c_float(drop_p),
                                      C.c_uint(drop_site), _p(rng), **wk)

    def gemm_nt_banded(self, A, Bw, Cm, M, N, segs, bands, bias, residual, drop_p, drop_site, rng):
        """csi_gemm_nt with per-segment output-column bands (the three Conv1d branches of an encoder as one GEMM)."""
        wk = self._work("gemm_nt", locals())
        if self._prof is not None:
            self._tag = f"M={M} N={N} K={sum(s[3] for s in segs)} nseg={len(segs)} banded out={Cm.dtype}"
        arr = (Seg * len(segs))(*[Seg(*s) for s in segs])
        barr = (Band * len(bands))(*[Band(*b) for b in bands])
        self._call("csi_gemm_nt_banded", _p(A), _ld(A), _p(Bw), _ld(Bw), _dt(A), _p(Cm), _ld(Cm), _dt(Cm), M, N, arr, barr,
                   len(segs), _p(bias), _p(residual), _ld(residual), C.c_float(drop_p), C.c_uint(drop_site), _p(rng), **wk)

    def tn_workspace(self):
        """Registers (once per device and stream) the workspace that makes csi_gemm_tn on the current stream reduce its token
        chunks in two stages (partial sums + one fixed-order reduce) instead of with fp32 atomics: faster and bit-reproducible."""
        if not TN_TWO_STAGE:
            return
        st = torch.cuda.current_stream(self.device).cuda_stream
        key = (self.device.index or 0, st)
        if key not in _TN_WS:
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            ws = torch.empty(sms * 128 * 512, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                rc = self.lib.csi_gemm_tn_workspace(C.c_void_p(st), _p(ws), C.c_longlong(ws.numel()))
            if rc != 0:
                raise RuntimeError(f"csi_gemm_tn_workspace failed ({rc}): {self.lib.csi_last_error().decode()}")
            _TN_WS[key] = ws

    def gemm_tn(self, A, Bv, Cm, ldc, c_col_stride, M, Na, segs, i_grp=NO_GRP, q_grp=NO_GRP):
        wk = self._work("gemm_tn", locals())
        if A.dtype == torch.bfloat16:
            self.tn_workspace()
        if self._prof is not None:
            self._tag = f"M={M} Na={Na} nlen={sum(s[3] for s in segs)} nseg={len(segs)} cs={c_col_stride}"
        arr = (SegTN * len(segs))(*[SegTN(*s) for s in segs])
        # the tcgen05 path with a registered workspace is two kernels (partial sums + reduce): count both
        two = A.dtype == torch.bfloat16 and TN_TWO_STAGE and M >= 64
        self._call("csi_gemm_tn", _p(A), _ld(A), _p(Bv), _ld(Bv), _dt(A), _p(Cm), ldc, c_col_stride, M, Na, arr,
                   len(segs), Grp(*i_grp), Grp(*q_grp), launches=2 if two else 1, **wk)

    def colsum_tokens(self, A, B, L, halo, ncols, out, grp=NO_GRP):
        wk = self._work("colsum_tokens", locals())
        self._call("csi_colsum_tokens", _p(A), _ld(A), _dt(A), B, L, halo, ncols, Grp(*grp), _p(out), **wk)

    def attn_fwd(self, qkv, o, lse, B, L, d, H, hp, halo):
        wk = self._work("attn_fwd", locals())
        self._call("csi_attn_fwd", _p(qkv), _ld(qkv), _p(o), _ld(o), _dt(qkv), _p(lse), B, L, d, H, hp, halo, **wk)

    def attn_bwd(self, qkv, o, dout, dqkv, lse, B, L, d, H, hp, halo, dbias=None):
        wk = self._work("attn_bwd", locals())
        self._call("csi_attn_bwd", _p(qkv), _ld(qkv), _p(o), _ld(o), _p(dout), _ld(dout), _p(dqkv), _ld(dqkv),
                                       _dt(qkv), _p(lse), B, L, d, H, hp, halo, _p(dbias), **wk)

    def bn_stats(self, z, B, L, halo, ncols, sums):
        wk = self._work("bn_stats", locals())
        self._call("csi_bn_stats", _p(z), _ld(z), _dt(z), B, L, halo, ncols, _p(sums), **wk)

    def bn_finalize(self, sums, Dp, d, nbr, count, conv_bias, run_mean, run_var, nbt, momentum, eps, mean, invstd):
        wk = self._work("bn_finalize", locals())
        self._call("csi_bn_finalize", _p(sums), Dp, d, nbr, C.c_longlong(count), self._ptr3(conv_bias),
                                          self._ptr3(run_mean), self._ptr3(run_var), self._ptr3(nbt),
                                          C.c_float(momentum), C.c_float(eps), _p(mean), _p(invstd), **wk)

    def bn_eval_prepare(self, Dp, d, nbr, conv_bias, run_mean, run_var, eps, mean, invstd):
        wk = self._work("bn_eval_prepare", locals())
        self._call("csi_bn_eval_prepare", Dp, d, nbr, self._ptr3(conv_bias), self._ptr3(run_mean),
                                              self._ptr3(run_var), C.c_float(eps), _p(mean), _p(invstd), **wk)

    def bn_act_fwd(self, z, mean, invstd, gamma, beta, t_res, out, B, L, d, halo, nbr, p_branch, site_branch, p_out,
                   site_out, rng, masks=None):
        wk = self._work("bn_act_fwd", locals())
        self._call("csi_bn_act_fwd", _p(z), _ld(z), _dt(z), _p(mean), _p(invstd), self._ptr3(gamma),
                                         self._ptr3(beta), _p(t_res), _ld(t_res), _p(out), _ld(out), B, L, d, halo,
                                         nbr, C.c_float(p_branch), C.c_uint(site_branch), C.c_float(p_out),
                                         C.c_uint(site_out), _p(rng), _p(masks), **wk)

    def bn_act_bwd_reduce(self, dout, z, mean, invstd, gamma, beta, B, L, d, halo, nbr, p_branch, site_branch, p_out,
                          site_out, rng, red, masks=None):
        wk = self._work("bn_act_bwd_reduce", locals())
        self._call("csi_bn_act_bwd_reduce", _p(dout), _ld(dout), _p(z), _ld(z), _dt(z), _p(mean), _p(invstd),
                                                self._ptr3(gamma), self._ptr3(beta), B, L, d, halo, nbr,
                                                C.c_float(p_branch), C.c_uint(site_branch), C.c_float(p_out),
                                                C.c_uint(site_out), _p(rng), _p(masks), _p(red), **wk)

    def bn_act_bwd_dz(self, dout, z, mean, invstd, gamma, beta, red, B, L, d, halo, nbr, p_branch, site_branch, p_out,
                      site_out, rng, dz, dgamma, dbeta, masks=None):
        wk = self._work("bn_act_bwd_dz", locals())
        self._call("csi_bn_act_bwd_dz", _p(dout), _ld(dout), _p(z), _ld(z), _dt(z), _p(mean), _p(invstd),
                                            self._ptr3(gamma), self._ptr3(beta), _p(red), B, L, d, halo, nbr,
                                            C.c_float(p_branch), C.c_uint(site_branch), C.c_float(p_out),
                                            C.c_uint(site_out), _p(rng), _p(masks), _p(dz), _ld(dz), self._ptr3(dgamma),
                                            self._ptr3(dbeta), **wk)

    def head_reduce_fwd(self, p, B, L, halo, N, n0, k0, k1, feat):
        wk = self._work("head_reduce_fwd", locals())
        self._call("csi_head_reduce_fwd", _p(p), _ld(p), _dt(p), B, L, halo, N, n0, k0, k1, _p(feat), _ld(feat), **wk)

    def head_reduce_bwd(self, dfeat, p, B, L, halo, N, n0, k0, k1, dp):
        wk = self._work("head_reduce_bwd", locals())
        self._call("csi_head_reduce_bwd", _p(dfeat), _ld(dfeat), _p(p), _ld(p), _dt(p), B, L, halo, N, n0, k0, k1,
                                              _p(dp), _ld(dp), **wk)

    def dropout_rows(self, inp, out, rows, cols, p, site, rng):
        wk = self._work("dropout_rows", locals())
        self._call("csi_dropout_rows", _p(inp), _ld(inp), _p(out), _ld(out), _dt(out), rows, cols, C.c_float(p),
                                           C.c_uint(site), _p(rng), **wk)

    def bce_logits(self, z, y, rows, cols, pos_weight, grad_scale, loss, dz):
        wk = self._work("bce_logits", locals())
        self._call("csi_bce_logits", _p(z), _ld(z), _p(y), _ld(y), rows, cols, C.c_float(pos_weight),
                                         C.c_float(grad_scale), _p(loss), _p(dz), _ld(dz), **wk)

    def smooth_l1(self, z, y, rows, cols, beta, grad_scale, loss, dz):
        self._call("csi_smooth_l1", _p(z), _ld(z), _p(y), _ld(y), rows, cols, C.c_float(beta), C.c_float(grad_scale),
                   _p(loss), _p(dz), _ld(dz))

    def perm_ce(self, z, y, B, heads, classes, cpitch, grad_scale, loss, dz, best_perm=None):
        self._call("csi_perm_ce", _p(z), _ld(z), _p(y), _ld(y), B, heads, classes, cpitch, C.c_float(grad_scale),
                   _p(loss), _p(dz), _ld(dz), _p(best_perm))

    def predict_counts(self, logits, rows, users, classes, threshold, counts):
        self._call("csi_predict_counts", _p(logits), _ld(logits), rows, users, classes, C.c_float(threshold), _p(counts))

    def adam_flat(self, p, g, m, v, n, lr, b1, b2, eps, wd, step, grad_scale):
        wk = self._work("adam_flat", locals())
        self._call("csi_adam_flat", _p(p), _p(g), _p(m), _p(v), C.c_longlong(n), C.c_float(lr), C.c_float(b1),
                                        C.c_float(b2), C.c_float(eps), C.c_float(wd), _p(step), C.c_float(grad_scale), **wk)

    def advance_counters(self, rng, step):
        wk = self._work("advance_counters", locals())
        self._call("csi_advance_counters", _p(rng), _p(step), **wk)

    def fill_f32(self, t, v):
        wk = self._work("fill_f32", locals())
        self._call("csi_fill_f32", _p(t), C.c_longlong(t.numel()), C.c_float(v), **wk)

    def fill_f64(self, t, v=0.0):
        self._call("csi_fill_f64", _p(t), C.c_longlong(t.numel()), C.c_double(v))

    def copy_f32(self, dst, src, n):
        self._call("csi_copy_f32", _p(dst), _p(src), C.c_longlong(n))

    # -------------------------------------------------------------- CSI-as-image path (cnn2d.cu)
    def nhwc_stats(self, x, rows, Cc, sums):
        self._call("csi_nhwc_stats", _p(x), _dt(x), C.c_longlong(rows), Cc, _p(sums), nbytes=rows * Cc * x.element_size())

    def bn2d_finalize(self, sums, Cc, count, gamma, beta, run_mean, run_var, nbt, momentum, eps, training, mean, invstd, scale, shift):
        self._call("csi_bn2d_finalize", _p(sums), Cc, C.c_longlong(count), _p(gamma), _p(beta), _p(run_mean), _p(run_var), _p(nbt),
                   C.c_float(momentum), C.c_float(eps), int(training), _p(mean), _p(invstd), _p(scale), _p(shift))

    def im2col_bn(self, x, B, H, W, Cc, k, s, scale, shift, col, Kp):
        OH, OW = (H - k) // s + 1, (W - k) // s + 1
        self._call("csi_im2col_bn", _p(x), _dt(x), B, H, W, Cc, k, s, _p(scale), _p(shift), _p(col), _dt(col), Kp,
                   nbytes=B * OH * OW * Kp * col.element_size() + B * H * W * Cc * x.element_size())

    def col2im(self, gcol, B, H, W, Cc, k, s, Kp, g):
        OH, OW = (H - k) // s + 1, (W - k) // s + 1
        self._call("csi_col2im", _p(gcol), _dt(gcol), B, H, W, Cc, k, s, Kp, _p(g),
                   nbytes=B * OH * OW * Kp * gcol.element_size() + B * H * W * Cc * 4)

    def act_drop_fwd(self, z, y, n, p, site, rng, mask):
        self._call("csi_act_drop_fwd", _p(z), _p(y), _dt(z), C.c_longlong(n), C.c_float(p), C.c_uint(site), _p(rng), _p(mask),
                   nbytes=2 * n * z.element_size())

    def bn2d_bwd_reduce(self, g, g_div, g_scale, x, rows, Cc, mean, invstd, sums):
        self._call("csi_bn2d_bwd_reduce", _p(g), _dt(g), C.c_longlong(g_div), C.c_float(g_scale), _p(x), _dt(x), C.c_longlong(rows), Cc, _p(mean),
                   _p(invstd), _p(sums), nbytes=rows * Cc * x.element_size() + (rows // g_div) * Cc * g.element_size())

    def bn2d_bwd_apply(self, g, g_div, g_scale, x, zprev, mask, drop_p, rows, Cc, mean, invstd, gamma, sums, gz, dgamma, dbeta):
        self._call("csi_bn2d_bwd_apply", _p(g), C.c_longlong(g_div), C.c_float(g_scale), _p(x), _p(zprev), _dt(x), _p(mask), C.c_float(drop_p),
                   C.c_longlong(rows), Cc, _p(mean), _p(invstd), _p(gamma), _p(sums), _p(gz), _p(dgamma), _p(dbeta),
                   nbytes=rows * Cc * 3 * x.element_size() + (rows // g_div) * Cc * 4)

    def pool_bn_fwd(self, y, B, P, Cc, scale, shift, feat, featd):
        self._call("csi_pool_bn_fwd", _p(y), _dt(y), B, P, Cc, _p(scale), _p(shift), _p(feat), _p(featd))

    def conv2d_pack(self, w, N, Cc, k, wf, Kp, wb, Np):
        self._call("csi_conv2d_pack", _p(w), N, Cc, k, _p(wf), Kp, _p(wb), Np, _dt(wf))

    def conv2d_unpack_grad(self, gs, N, Cc, k, Kp, gw):
        self._call("csi_conv2d_unpack_grad", _p(gs), N, Cc, k, Kp, _p(gw))

    def bn0_grads(self, sums, wf, ldw, N, K, bias, gamma0, beta0, dgamma0, dbeta0, dbias):
        self._call("csi_bn0_grads", _p(sums), _p(wf), ldw, _dt(wf), N, K, _p(bias), _p(gamma0), _p(beta0), _p(dgamma0), _p(dbeta0),
                   _p(dbias))

    def gather_aug(self, x, offs, lens, B, T, F, out, augment, rng):
        self._call("csi_gather_aug", _p(x), _p(offs), _p(lens), B, T, F, _p(out), int(augment), _p(rng), nbytes=2 * B * T * F * 4)
