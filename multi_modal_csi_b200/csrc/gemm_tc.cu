// tcgen05 / TMEM / TMA contraction kernels (placeholder until the kernel lands: reports "not eligible").
#include "common.cuh"
extern "C" int csi_gemm_nt_tc_ok(int lda, int ldb, int ldc, int M, int N, const csi_seg* segs, int nseg) { return 0; }
extern "C" int csi_gemm_nt_tc(const void*, int, const void*, int, void*, int, int, int, int, const csi_seg*, int,
                              const float*, const float*, int, float, unsigned, const unsigned long long*, void*) {
    csi_set_error("csi_gemm_nt_tc: not built");
    return CSI_ERR_ARG;
}
