// Entry points whose implementation depends on operand type / shape.
//   bf16 operands -> the tcgen05 kernels (gemm_tc3.cu, gemm_tn3.cu, attention_tc.cu)
//   fp32 operands -> the FFMA "fp32 parity" kernels (gemm_simt.cu, attention.cu)
// A bf16 call whose shape the tcgen05 kernel cannot take (fewer than 64 token rows, unaligned pitches: only the toy
// shapes of the unit tests) still runs on an FFMA kernel, but never silently: every such call is counted
// (csi_dispatch_counts) and, in strict mode (csi_set_strict_tc(1): bench.py and the full-size tests), it is an error.
#include "tc_common.cuh"
#include <stdlib.h>
#include <stdint.h>

extern "C" int csi_gemm_nt_simt(const void*, int, const void*, int, int, void*, int, int, int, int, const csi_seg*, int,
                                const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_tn_simt(const void*, int, const void*, int, int, float*, int, int, int, int, const csi_seg_tn*,
                                int, csi_grp, csi_grp, void*);
extern "C" int csi_attn_fwd_simt(const void*, int, void*, int, int, float*, int, int, int, int, int, int, void*);
extern "C" int csi_attn_bwd_simt(const void*, int, const void*, int, const void*, int, void*, int, int, const float*, int,
                                 int, int, int, int, int, void*);
extern "C" int csi_gemm_nt_tc3(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, int,
                               const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_nt_tc3_banded(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, const csi_band*, int,
                                      const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_tn_tc3(const void*, int, const void*, int, float*, int, int, int, int, const csi_seg_tn*, int, csi_grp,
                               csi_grp, void*);
extern "C" int csi_attn_mma_ok(int L, int d, int H, int hp);
extern "C" int csi_attn_fwd_mma(const void*, int, void*, int, float*, int, int, int, int, int, int, void*);
extern "C" int csi_attn_bwd_mma(const void*, int, const void*, int, const void*, int, void*, int, const float*, int, int, int,
                                int, int, int, float*, void*);
extern "C" int csi_attn_tc_ok(int L, int d, int H, int hp);
extern "C" int csi_attn_bwd_tc_ok(int L, int d, int H, int hp);
extern "C" int csi_attn_fwd_tc(const void*, int, void*, int, float*, int, int, int, int, int, int, void*);
extern "C" int csi_attn_bwd_tc(const void*, int, const void*, int, const void*, int, void*, int, const float*, int, int, int,
                               int, int, int, float*, void*);

static int g_force_simt = -1;
static bool force_simt() {
    if (g_force_simt < 0) { const char* e = getenv("CSI_FORCE_SIMT"); g_force_simt = (e && e[0] == '1') ? 1 : 0; }
    return g_force_simt == 1;
}
// tests flip this at run time (1 = FFMA kernels only, 0 = tensor-core kernels where eligible)
extern "C" int csi_set_force_simt(int on) { g_force_simt = on ? 1 : 0; return CSI_OK; }

// ---- bf16 dispatch accounting: [0] = calls served by a tcgen05 kernel, [1] = bf16 calls that ran on an FFMA kernel,
//      [2] = attention calls served by the mma.sync tensor-core kernel
static long long g_counts[3] = {0, 0, 0};
static int g_strict = 0;
extern "C" int csi_set_strict_tc(int on) { g_strict = on ? 1 : 0; return CSI_OK; }
extern "C" int csi_dispatch_counts(long long* tc_calls, long long* fallback_calls, long long* mma_sync_calls, int reset) {
    if (tc_calls) *tc_calls = g_counts[0];
    if (fallback_calls) *fallback_calls = g_counts[1];
    if (mma_sync_calls) *mma_sync_calls = g_counts[2];
    if (reset) g_counts[0] = g_counts[1] = g_counts[2] = 0;
    return CSI_OK;
}
#define CSI_FALLBACK(what)                                                                                            \
    do {                                                                                                              \
        if (!force_simt()) {                                                                                          \
            ++g_counts[1];                                                                                            \
            if (g_strict) { csi_set_error("%s: bf16 shape not eligible for the tcgen05 kernel (strict mode)", what); return CSI_ERR_ARG; } \
        }                                                                                                             \
    } while (0)

extern "C" int csi_gemm_nt_tc_ok(int lda, int ldb, int ldc, int M, int N, const csi_seg* segs, int nseg) {
    if (M < 1 || N < 1 || nseg < 1 || nseg > CSI_MAX_SEGS) return 0;
    if (lda % 8 || ldb % 8 || ldc % 2) return 0;               // 16-byte TMA row pitch; paired epilogue stores
    for (int i = 0; i < nseg; ++i)
        if (segs[i].klen % 16 || segs[i].klen <= 0 || segs[i].a_col_off % 8 || segs[i].b_col_off % 8) return 0;
    return get_encode() != nullptr;
}

extern "C" int csi_gemm_tn_tc_ok(int lda, int ldb, int M, int Na, const csi_seg_tn* segs, int nseg) {
    if (M < 64 || Na < 1 || nseg < 1 || nseg > CSI_MAX_SEGS) return 0;
    if (lda % 8 || ldb % 8) return 0;
    for (int i = 0; i < nseg; ++i)
        if (segs[i].nlen <= 0 || segs[i].b_col_off % 8) return 0;
    return get_encode() != nullptr;
}

extern "C" int csi_gemm_nt(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc, int c_dtype,
                           int M, int N, const csi_seg* segs, int nseg, const float* bias, const float* residual,
                           int ldr, float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    return csi_gemm_nt_banded(A, lda, Bw, ldb, ab_dtype, C, ldc, c_dtype, M, N, segs, nullptr, nseg, bias, residual, ldr, drop_p,
                              drop_site, rng, stream);
}

extern "C" int csi_gemm_nt_banded(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc, int c_dtype,
                                  int M, int N, const csi_seg* segs, const csi_band* bands, int nseg, const float* bias,
                                  const float* residual, int ldr, float drop_p, unsigned drop_site,
                                  const unsigned long long* rng, void* stream) {
    if (ab_dtype == CSI_BF16) {
        const int es = c_dtype == CSI_BF16 ? 2 : 4;
        const bool aligned = ((long long)ldc * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                             (!residual || (c_dtype != CSI_BF16 && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0));
        if (!force_simt() && aligned && csi_gemm_nt_tc_ok(lda, ldb, ldc, M, N, segs, nseg)) {
            ++g_counts[0];
            return csi_gemm_nt_tc3_banded(A, lda, Bw, ldb, C, ldc, c_dtype, M, N, segs, bands, nseg, bias, residual, ldr, drop_p,
                                          drop_site, rng, stream);
        }
        CSI_FALLBACK("csi_gemm_nt");
    }
    return csi_gemm_nt_simt(A, lda, Bw, ldb, ab_dtype, C, ldc, c_dtype, M, N, segs, nseg, bias, residual, ldr, drop_p,
                            drop_site, rng, stream);
}

extern "C" int csi_gemm_tn(const void* A, int lda, const void* Bv, int ldb, int ab_dtype, float* C, int ldc,
                           int c_col_stride, int M, int Na, const csi_seg_tn* segs, int nseg, csi_grp i_grp, csi_grp q_grp,
                           void* stream) {
    if (ab_dtype == CSI_BF16) {
        if (!force_simt() && csi_gemm_tn_tc_ok(lda, ldb, M, Na, segs, nseg)) {
            ++g_counts[0];
            return csi_gemm_tn_tc3(A, lda, Bv, ldb, C, ldc, c_col_stride, M, Na, segs, nseg, i_grp, q_grp, stream);
        }
        CSI_FALLBACK("csi_gemm_tn");
    }
    return csi_gemm_tn_simt(A, lda, Bv, ldb, ab_dtype, C, ldc, c_col_stride, M, Na, segs, nseg, i_grp, q_grp, stream);
}

// attention core.  Two tensor-core implementations exist for bf16: the tcgen05/TMEM kernels (attention_tc.cu,
// attention_tc_bwd.cu) and the mma.sync kernels (attention_mma.cu).  With head widths of 15-54 channels a tcgen05 tile is
// mostly padding and the softmax (MUFU + TMEM round trips at 3 warps per scheduler) bounds the kernel, so which one is
// faster depends on the shape; measured on B200 at B=256 (scripts/test_attn_tc.py, profiles/r2_attention_tc_vs_mma.txt):
//     (L, d, hp)        fwd tc / mma      bwd tc / mma
//     (150, 270, 32)     72 /  53 us      239 / 152 us
//     (270, 150, 16)    242 /  84 us      521 / 214 us
//     (150, 540, 64)     84 /  98 us      372 / 336 us
// mode 0 (default) picks the measured-best kernel per shape, 1 = tcgen05 wherever eligible, 2 = mma.sync only.
// CSI_ATTN_IMPL=<mode> or csi_set_attn_impl(fwd_mode, bwd_mode).
static int g_attn_mode[2] = {-1, -1};
static int attn_mode(int which) {
    if (g_attn_mode[which] < 0) {
        const char* e = getenv("CSI_ATTN_IMPL");
        g_attn_mode[which] = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
    }
    return g_attn_mode[which];
}
extern "C" int csi_set_attn_impl(int fwd_mode, int bwd_mode) {
    g_attn_mode[0] = fwd_mode;
    g_attn_mode[1] = bwd_mode;
    return CSI_OK;
}
static bool attn_use_tc(int which, int L, int d, int H, int hp) {
    const int mode = attn_mode(which);
    if (mode == 2) return false;
    const bool ok = which == 0 ? csi_attn_tc_ok(L, d, H, hp) : csi_attn_bwd_tc_ok(L, d, H, hp);
    if (!ok) return false;
    if (mode == 1 || !csi_attn_mma_ok(L, d, H, hp)) return true;
    return which == 0 && hp == 64;                      // the measured-best table above
}

extern "C" int csi_attn_fwd(const void* qkv, int ld3, void* o, int ldo, int dtype, float* lse, int B, int L, int d,
                            int H, int hp, int halo, void* stream) {
    if (dtype == CSI_BF16 && !force_simt()) {
        if (attn_use_tc(0, L, d, H, hp)) {
            ++g_counts[0];
            return csi_attn_fwd_tc(qkv, ld3, o, ldo, lse, B, L, d, H, hp, halo, stream);
        }
        if (csi_attn_mma_ok(L, d, H, hp)) {
            ++g_counts[2];
            return csi_attn_fwd_mma(qkv, ld3, o, ldo, lse, B, L, d, H, hp, halo, stream);
        }
    }
    if (dtype == CSI_BF16) CSI_FALLBACK("csi_attn_fwd");
    return csi_attn_fwd_simt(qkv, ld3, o, ldo, dtype, lse, B, L, d, H, hp, halo, stream);
}

extern "C" int csi_attn_bwd(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo, void* dqkv,
                            int lddqkv, int dtype, const float* lse, int B, int L, int d, int H, int hp, int halo, float* dbias,
                            void* stream) {
    if (dtype == CSI_BF16 && !force_simt()) {
        if (attn_use_tc(1, L, d, H, hp)) {
            ++g_counts[0];
            return csi_attn_bwd_tc(qkv, ld3, o, ldo, dout, lddo, dqkv, lddqkv, lse, B, L, d, H, hp, halo, dbias, stream);
        }
        if (csi_attn_mma_ok(L, d, H, hp)) {
            ++g_counts[2];
            return csi_attn_bwd_mma(qkv, ld3, o, ldo, dout, lddo, dqkv, lddqkv, lse, B, L, d, H, hp, halo, dbias, stream);
        }
    }
    if (dtype == CSI_BF16) CSI_FALLBACK("csi_attn_bwd");
    int rc = csi_attn_bwd_simt(qkv, ld3, o, ldo, dout, lddo, dqkv, lddqkv, dtype, lse, B, L, d, H, hp, halo, stream);
    if (rc || !dbias) return rc;
    // the FFMA path has no fused bias gradient: column sums of dqkv over the valid tokens, compact (un-padded) index
    return csi_colsum_tokens(dqkv, lddqkv, dtype, B, L, halo, 3 * H * hp, csi_grp{d / H, hp}, dbias, stream);
}
