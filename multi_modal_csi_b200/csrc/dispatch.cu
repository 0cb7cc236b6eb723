// Entry points whose implementation depends on operand type / shape: bf16 contractions go to the tcgen05 kernels
// (gemm_tc.cu) when the shape qualifies, everything else to the FFMA kernels (gemm_simt.cu).
#include "common.cuh"
#include <stdlib.h>
#include <stdint.h>

extern "C" int csi_gemm_nt_simt(const void*, int, const void*, int, int, void*, int, int, int, int, const csi_seg*, int,
                                const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_tn_simt(const void*, int, const void*, int, int, float*, int, int, int, int, const csi_seg_tn*,
                                int, csi_grp, csi_grp, void*);
extern "C" int csi_attn_fwd_simt(const void*, int, void*, int, int, float*, int, int, int, int, int, int, void*);
extern "C" int csi_attn_bwd_simt(const void*, int, const void*, int, const void*, int, void*, int, int, const float*, int,
                                 int, int, int, int, int, void*);
extern "C" int csi_gemm_nt_tc(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, int,
                              const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_nt_tc_ok(int lda, int ldb, int ldc, int M, int N, const csi_seg* segs, int nseg);
extern "C" int csi_gemm_nt_tc2(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, int,
                               const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
extern "C" int csi_gemm_nt_tc3(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, int,
                               const float*, const float*, int, float, unsigned, const unsigned long long*, void*);
static int g_gemm_v2 = -1;
static bool gemm_v2() {
    if (g_gemm_v2 < 0) { const char* e = getenv("CSI_GEMM_V2"); g_gemm_v2 = (e && e[0] == '1') ? 1 : 0; }
    return g_gemm_v2 == 1;
}
extern "C" int csi_set_gemm_v2(int on) { g_gemm_v2 = on ? 1 : 0; return CSI_OK; }
static int g_gemm_v1 = -1;
static bool gemm_v1() {
    if (g_gemm_v1 < 0) { const char* e = getenv("CSI_GEMM_V1"); g_gemm_v1 = (e && e[0] == '1') ? 1 : 0; }
    return g_gemm_v1 == 1;
}
extern "C" int csi_gemm_tn_tc(const void*, int, const void*, int, float*, int, int, int, int, const csi_seg_tn*, int, csi_grp,
                              csi_grp, void*);
extern "C" int csi_gemm_tn_tc_ok(int lda, int ldb, int M, int Na, const csi_seg_tn* segs, int nseg);
extern "C" int csi_gemm_tn_tc3(const void*, int, const void*, int, float*, int, int, int, int, const csi_seg_tn*, int, csi_grp,
                               csi_grp, void*);
static int g_tn_v1 = -1;
static bool tn_v1() {
    if (g_tn_v1 < 0) { const char* e = getenv("CSI_GEMM_TN_V1"); g_tn_v1 = (e && e[0] == '1') ? 1 : 0; }
    return g_tn_v1 == 1;
}
extern "C" int csi_set_gemm_tn_v1(int on) { g_tn_v1 = on ? 1 : 0; return CSI_OK; }

extern "C" int csi_attn_mma_ok(int L, int d, int H, int hp);
extern "C" int csi_attn_fwd_mma(const void*, int, void*, int, float*, int, int, int, int, int, int, void*);
extern "C" int csi_attn_bwd_mma(const void*, int, const void*, int, const void*, int, void*, int, const float*, int, int, int,
                                int, int, int, float*, void*);

static int g_force_simt = -1;
static bool force_simt() {
    if (g_force_simt < 0) { const char* e = getenv("CSI_FORCE_SIMT"); g_force_simt = (e && e[0] == '1') ? 1 : 0; }
    return g_force_simt == 1;
}
// tests flip this at run time (1 = FFMA kernels only, 0 = tensor-core kernels where eligible)
extern "C" int csi_set_force_simt(int on) { g_force_simt = on ? 1 : 0; return CSI_OK; }

extern "C" int csi_gemm_nt(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc, int c_dtype,
                           int M, int N, const csi_seg* segs, int nseg, const float* bias, const float* residual,
                           int ldr, float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    const int es = c_dtype == CSI_BF16 ? 2 : 4;
    const bool v2_ok = ((long long)ldc * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                       (!residual || (c_dtype != CSI_BF16 && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0));
    if (ab_dtype == CSI_BF16 && !force_simt() && !gemm_v1() && !gemm_v2() && v2_ok && csi_gemm_nt_tc_ok(lda, ldb, ldc, M, N, segs, nseg))
        return csi_gemm_nt_tc3(A, lda, Bw, ldb, C, ldc, c_dtype, M, N, segs, nseg, bias, residual, ldr, drop_p, drop_site,
                               rng, stream);
    if (ab_dtype == CSI_BF16 && !force_simt() && !gemm_v1() && v2_ok && csi_gemm_nt_tc_ok(lda, ldb, ldc, M, N, segs, nseg))
        return csi_gemm_nt_tc2(A, lda, Bw, ldb, C, ldc, c_dtype, M, N, segs, nseg, bias, residual, ldr, drop_p, drop_site,
                               rng, stream);
    if (ab_dtype == CSI_BF16 && !force_simt() && csi_gemm_nt_tc_ok(lda, ldb, ldc, M, N, segs, nseg))
        return csi_gemm_nt_tc(A, lda, Bw, ldb, C, ldc, c_dtype, M, N, segs, nseg, bias, residual, ldr, drop_p,
                              drop_site, rng, stream);
    return csi_gemm_nt_simt(A, lda, Bw, ldb, ab_dtype, C, ldc, c_dtype, M, N, segs, nseg, bias, residual, ldr, drop_p,
                            drop_site, rng, stream);
}

extern "C" int csi_gemm_tn(const void* A, int lda, const void* Bv, int ldb, int ab_dtype, float* C, int ldc,
                           int c_col_stride, int M, int Na, const csi_seg_tn* segs, int nseg, csi_grp i_grp, csi_grp q_grp,
                           void* stream) {
    if (ab_dtype == CSI_BF16 && !force_simt() && !tn_v1() && csi_gemm_tn_tc_ok(lda, ldb, M, Na, segs, nseg))
        return csi_gemm_tn_tc3(A, lda, Bv, ldb, C, ldc, c_col_stride, M, Na, segs, nseg, i_grp, q_grp, stream);
    if (ab_dtype == CSI_BF16 && !force_simt() && csi_gemm_tn_tc_ok(lda, ldb, M, Na, segs, nseg))
        return csi_gemm_tn_tc(A, lda, Bv, ldb, C, ldc, c_col_stride, M, Na, segs, nseg, i_grp, q_grp, stream);
    return csi_gemm_tn_simt(A, lda, Bv, ldb, ab_dtype, C, ldc, c_col_stride, M, Na, segs, nseg, i_grp, q_grp, stream);
}

extern "C" int csi_attn_fwd(const void* qkv, int ld3, void* o, int ldo, int dtype, float* lse, int B, int L, int d,
                            int H, int hp, int halo, void* stream) {
    if (dtype == CSI_BF16 && !force_simt() && csi_attn_mma_ok(L, d, H, hp))
        return csi_attn_fwd_mma(qkv, ld3, o, ldo, lse, B, L, d, H, hp, halo, stream);
    return csi_attn_fwd_simt(qkv, ld3, o, ldo, dtype, lse, B, L, d, H, hp, halo, stream);
}

extern "C" int csi_attn_bwd(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo, void* dqkv,
                            int lddqkv, int dtype, const float* lse, int B, int L, int d, int H, int hp, int halo, float* dbias,
                            void* stream) {
    if (dtype == CSI_BF16 && !force_simt() && csi_attn_mma_ok(L, d, H, hp))
        return csi_attn_bwd_mma(qkv, ld3, o, ldo, dout, lddo, dqkv, lddqkv, lse, B, L, d, H, hp, halo, dbias, stream);
    int rc = csi_attn_bwd_simt(qkv, ld3, o, ldo, dout, lddo, dqkv, lddqkv, dtype, lse, B, L, d, H, hp, halo, stream);
    if (rc || !dbias) return rc;
    // the FFMA path has no fused bias gradient: column sums of dqkv over the valid tokens, compact (un-padded) index
    return csi_colsum_tokens(dqkv, lddqkv, dtype, B, L, halo, 3 * H * hp, csi_grp{d / H, hp}, dbias, stream);
}
