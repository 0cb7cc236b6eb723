// FFMA GEMMs: the fp32 "parity mode" contraction kernels (CSI_F32 operands) and the fall-through for shapes the
// tcgen05 path does not take (tiny N, unaligned leading dimensions).  bf16 operands are accepted too (converted to
// fp32 in shared memory) so that every tensor-core kernel has a same-input FFMA reference on the GPU.
#include "common.cuh"

#define ST(s) ((cudaStream_t)(s))

struct SegList { csi_seg s[CSI_MAX_SEGS]; int n; };
struct SegListTN { csi_seg_tn s[CSI_MAX_SEGS]; int n; };

// ------------------------------------------------------------------------------------------------ NT
// C[m,n] = sum_seg sum_q A[(m+shift)*lda + aoff + q] * B[n*ldb + boff + q]  -> epilogue
// 128x128 tile, BK=16, 256 threads, 8x8 micro tile.
#define NT_BM 128
#define NT_BN 128
#define NT_BK 16

template <typename T> __device__ __forceinline__ void load8(const T* p, float* out);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* out) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float* out) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}

template <typename TA, typename TC>
__global__ void __launch_bounds__(256) gemm_nt_kernel(
    const TA* __restrict__ A, int lda, const TA* __restrict__ Bw, int ldb, TC* __restrict__ C, int ldc, int M, int N,
    SegList segs, const float* __restrict__ bias, const float* __restrict__ residual, int ldr, float drop_p,
    unsigned drop_site, const unsigned long long* __restrict__ rng) {
    __shared__ float As[NT_BK][NT_BM + 4];
    __shared__ float Bs[NT_BK][NT_BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * NT_BM, n0 = blockIdx.x * NT_BN;
    const int lr = tid >> 1, lk = (tid & 1) * 8;          // loader: row lr of the tile, 8 consecutive k
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int s = 0; s < segs.n; ++s) {
        const csi_seg sg = segs.s[s];
        const bool a_ok = (m0 + lr) < M, b_ok = (n0 + lr) < N;
        const TA* ap = A + ((long long)(m0 + lr) + sg.a_row_shift) * lda + sg.a_col_off + lk;
        const TA* bp = Bw + (long long)(n0 + lr) * ldb + sg.b_col_off + lk;
        for (int k0 = 0; k0 < sg.klen; k0 += NT_BK) {
            float av[8], bv[8];
            if (a_ok) load8<TA>(ap + k0, av); else { for (int i = 0; i < 8; ++i) av[i] = 0.f; }
            if (b_ok) load8<TA>(bp + k0, bv); else { for (int i = 0; i < 8; ++i) bv[i] = 0.f; }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) { As[lk + i][lr] = av[i]; Bs[lk + i][lr] = bv[i]; }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < NT_BK; ++kk) {
                float a[8], b[8];
                float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
    const bool drop = drop_p > 0.f;
    DropCtx dc;
    if (drop) dc = drop_ctx(rng, drop_p);
    const int ld8 = ((N + 15) & ~15) >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (drop) v *= drop_scale1(dc, drop_site, (unsigned long long)m, ld8, n);
            if (residual) v += residual[(long long)m * ldr + n];
            stf<TC>(C + (long long)m * ldc + n, v);
        }
    }
}

extern "C" int csi_gemm_nt_simt(const void* A, int lda, const void* Bw, int ldb, int ab_dtype, void* C, int ldc,
                                int c_dtype, int M, int N, const csi_seg* segs, int nseg, const float* bias,
                                const float* residual, int ldr, float drop_p, unsigned drop_site,
                                const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(A && Bw && C && segs, "null pointer");
    CSI_CHECK_ARG(nseg >= 1 && nseg <= CSI_MAX_SEGS, "1..32 segments");
    CSI_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "lda/ldb must be multiples of 8 elements");
    CSI_CHECK_ARG(!(drop_p > 0.f) || rng, "dropout needs rng");
    if (M == 0 || N == 0) return CSI_OK;
    SegList sl;
    sl.n = nseg;
    for (int i = 0; i < nseg; ++i) {
        sl.s[i] = segs[i];
        CSI_CHECK_ARG(segs[i].klen % 16 == 0 && segs[i].a_col_off % 8 == 0 && segs[i].b_col_off % 8 == 0,
                      "segment klen must be a multiple of 16 and offsets multiples of 8");
    }
    dim3 grid(cdiv(N, NT_BN), cdiv(M, NT_BM));
#define GO(TA, TC) gemm_nt_kernel<TA, TC><<<grid, 256, 0, ST(stream)>>>((const TA*)A, lda, (const TA*)Bw, ldb, (TC*)C, ldc, \
        M, N, sl, bias, residual, ldr, drop_p, drop_site, rng)
    const bool ab = ab_dtype == CSI_BF16, cb = c_dtype == CSI_BF16;
    if (ab && cb) GO(bf16, bf16); else if (ab) GO(bf16, float); else if (cb) GO(float, bf16); else GO(float, float);
#undef GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ TN (weight gradients)
// C[i*ldc + coff + q*cs] += sum_{m in chunk} A[m*lda + i] * Bv[(m+shift)*ldb + boff + q]
// 64x64 tile, 16 rows of m per step, 256 threads (4x4 micro tile), split over m chunks with atomics.
#define TN_BI 64
#define TN_BQ 64
#define TN_BK 16
#define TN_CHUNK 1024

template <typename TA>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const TA* __restrict__ A, int lda, const TA* __restrict__ Bv,
                                                      int ldb, float* __restrict__ C, int ldc, int cs, int M, int Na,
                                                      SegListTN segs, int qtiles_per_seg, csi_grp ig, csi_grp qg) {
    __shared__ float As[TN_BK][TN_BI + 4];
    __shared__ float Bs[TN_BK][TN_BQ + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int seg_i = blockIdx.y / qtiles_per_seg, qt = blockIdx.y % qtiles_per_seg;
    const csi_seg_tn sg = segs.s[seg_i];
    const int i0 = blockIdx.x * TN_BI, q0 = qt * TN_BQ;
    if (q0 >= sg.nlen) return;
    const int mbeg = blockIdx.z * TN_CHUNK, mend = min(M, mbeg + TN_CHUNK);
    const int lk = tid >> 4, lc = (tid & 15) * 4;           // loader: row lk (of 16), 4 consecutive columns
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int mk = mbeg; mk < mend; mk += TN_BK) {
        const int m = mk + lk;
        float av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + lc + u, q = q0 + lc + u;
            av[u] = (m < mend && i < Na) ? ldv<TA>(A + (long long)m * lda + i) : 0.f;
            bv[u] = (m < mend && q < sg.nlen) ? ldv<TA>(Bv + ((long long)m + sg.b_row_shift) * ldb + sg.b_col_off + q) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) { As[lk][lc + u] = av[u]; Bs[lk][lc + u] = bv[u]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TN_BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ii = i0 + ty * 4 + i;
        if (ii >= Na) continue;
        const int ic = grp_to_compact(ii, ig);
        if (ic < 0) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = q0 + tx * 4 + j;
            if (q >= sg.nlen) continue;
            const int qc = grp_to_compact(q, qg);
            if (qc < 0) continue;
            atomicAdd(C + (long long)ic * ldc + sg.c_off + (long long)qc * cs, acc[i][j]);
        }
    }
}

extern "C" int csi_gemm_tn_simt(const void* A, int lda, const void* Bv, int ldb, int ab_dtype, float* C, int ldc,
                                int c_col_stride, int M, int Na, const csi_seg_tn* segs, int nseg, csi_grp ig, csi_grp qg,
                                void* stream) {
    CSI_CHECK_ARG(A && Bv && C && segs, "null pointer");
    CSI_CHECK_ARG(nseg >= 1 && nseg <= CSI_MAX_SEGS, "1..32 segments");
    if (M == 0 || Na == 0) return CSI_OK;
    SegListTN sl;
    sl.n = nseg;
    int maxn = 0;
    for (int i = 0; i < nseg; ++i) { sl.s[i] = segs[i]; if (segs[i].nlen > maxn) maxn = segs[i].nlen; }
    const int qt = cdiv(maxn, TN_BQ);
    dim3 grid(cdiv(Na, TN_BI), qt * nseg, cdiv(M, TN_CHUNK));
    if (ab_dtype == CSI_BF16)
        gemm_tn_kernel<bf16><<<grid, 256, 0, ST(stream)>>>((const bf16*)A, lda, (const bf16*)Bv, ldb, C, ldc, c_col_stride, M, Na, sl, qt, ig, qg);
    else
        gemm_tn_kernel<float><<<grid, 256, 0, ST(stream)>>>((const float*)A, lda, (const float*)Bv, ldb, C, ldc, c_col_stride, M, Na, sl, qt, ig, qg);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
