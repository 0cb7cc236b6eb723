// tcgen05 / TMEM / TMA contraction kernels for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   csi_gemm_nt_tc : C[m,n] = sum_seg A[(m+shift)*lda + aoff + q] * B[n*ldb + boff + q]  (+bias, dropout, +residual)
//                    Linear layers, Conv1d forward and Conv1d data-gradient of the THAT encoder: every convolution
//                    tap is a K-segment whose A tile is fetched by TMA at a row-shifted coordinate of the same
//                    token buffer, so no im2col matrix ever exists.
//
// Structure (one 128 x BN output tile per CTA, 2 CTAs resident per SM so one CTA's epilogue overlaps the other's
// main loop):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d into a 128B-swizzled K-major smem ring (mbarrier tx-count)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16, fp32 accumulate)
//   warps 2-5: epilogue -- tcgen05.ld 32x32b, bias / dropout / residual, direct vectorised global stores
#include "tc_common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define TC_STAGES 3
#define TC_THREADS 192

struct SegList { csi_seg s[CSI_MAX_SEGS]; int n; };

struct NtParams {
    void* C; int ldc; int M, N, BN;
    const float* bias; const float* residual; int ldr;
    float drop_p; unsigned drop_site; const unsigned long long* rng;
    int row_base;                 // added to (m0 + shift) to form the TMA row coordinate (tensor map starts at the lowest row read)
    uint32_t tmem_cols;
};

template <typename TC>
__global__ void __launch_bounds__(TC_THREADS) gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, NtParams p,
                                                                SegList segs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * p.BN;
    const uint32_t a_bytes = TC_BM * TC_BK * 2, b_bytes = (uint32_t)p.BN * TC_BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int s = 0; s < segs.n; ++s) {
                const csi_seg sg = segs.s[s];
                for (int k0 = 0; k0 < sg.klen; k0 += TC_BK, ++it) {
                    const int stage = it % TC_STAGES;
                    const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(&empty_bar[stage], ph ^ 1u);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], stage_bytes);
                    tma_load_2d(&tmA, &full_bar[stage], sa, sg.a_col_off + k0, m0 + sg.a_row_shift + p.row_base);
                    tma_load_2d(&tmB, &full_bar[stage], sa + a_bytes, sg.b_col_off + k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(TC_BM, p.BN);
            int it = 0;
            uint32_t first = 1;
            for (int s = 0; s < segs.n; ++s) {
                const csi_seg sg = segs.s[s];
                for (int k0 = 0; k0 < sg.klen; k0 += TC_BK, ++it) {
                    const int stage = it % TC_STAGES;
                    const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(&full_bar[stage], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_kmajor_desc(sa), bdesc = make_kmajor_desc(sa + a_bytes);
                    const int ksteps = min(TC_BK, sg.klen - k0) >> 4;
                    for (int k = 0; k < ksteps; ++k) {
                        // +32 B per 16-element K step inside the 128 B swizzle row: start-address field += 2
                        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                        first = 0;
                    }
                    umma_commit(&empty_bar[stage]);        // frees the smem slot once these MMAs have read it
                }
            }
            umma_commit(&tmem_full_bar);                   // accumulator complete
        }
    } else {
        // ---------------- epilogue: TMEM lane quarter (warp % 4) -> 32 rows of the tile, one row per thread
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const bool drop = p.drop_p > 0.f;
        DropCtx dc;
        if (drop) dc = drop_ctx(p.rng, p.drop_p);
        const int ld8 = ((p.N + 15) & ~15) >> 3;
        TC* crow = reinterpret_cast<TC*>(p.C) + (size_t)m * p.ldc;
        const float* rrow = p.residual ? p.residual + (size_t)m * p.ldr : nullptr;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (m < p.M) {
                const int nb = n0 + c0;
                float ks[16];
                if (drop) {
                    float k0[8], k1[8];
                    drop_scales8(dc, p.drop_site, (unsigned long long)m * ld8 + (nb >> 3), k0);
                    drop_scales8(dc, p.drop_site, (unsigned long long)m * ld8 + (nb >> 3) + 1, k1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { ks[j] = k0[j]; ks[8 + j] = k1[j]; }
                }
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const int n = nb + j;
                    if (n >= p.N) break;
                    float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
                    const bool two = (n + 1) < p.N;
                    if (p.bias) { v0 += p.bias[n]; if (two) v1 += p.bias[n + 1]; }
                    if (drop) { v0 *= ks[j]; v1 *= ks[j + 1]; }
                    if (rrow) { v0 += rrow[n]; if (two) v1 += rrow[n + 1]; }
                    if (two) st2<TC>(crow + n, make_float2(v0, v1));
                    else stf<TC>(crow + n, v0);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int pick_bn(int N) {
    const int tiles = (N + 127) / 128;
    int bn = ((N + tiles - 1) / tiles + 15) & ~15;
    if (bn < 16) bn = 16;
    return bn;
}

extern "C" int csi_gemm_nt_tc_ok(int lda, int ldb, int ldc, int M, int N, const csi_seg* segs, int nseg) {
    if (M < 1 || N < 1 || nseg < 1 || nseg > CSI_MAX_SEGS) return 0;
    if (lda % 8 || ldb % 8 || ldc % 2) return 0;               // 16-byte TMA row pitch; paired epilogue stores
    for (int i = 0; i < nseg; ++i)
        if (segs[i].klen % 16 || segs[i].klen <= 0 || segs[i].a_col_off % 8 || segs[i].b_col_off % 8) return 0;
    return get_encode() != nullptr;
}

extern "C" int csi_gemm_nt_tc(const void* A, int lda, const void* Bw, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                              const csi_seg* segs, int nseg, const float* bias, const float* residual, int ldr,
                              float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(A && Bw && C && segs, "null pointer");
    CSI_CHECK_ARG(csi_gemm_nt_tc_ok(lda, ldb, ldc, M, N, segs, nseg), "shape not eligible for the tcgen05 kernel");
    CSI_CHECK_ARG(!(drop_p > 0.f) || rng, "dropout needs rng");
    SegList sl;
    sl.n = nseg;
    int min_shift = 0, max_shift = 0, a_cols = 0, b_cols = 0;
    for (int i = 0; i < nseg; ++i) {
        sl.s[i] = segs[i];
        if (segs[i].a_row_shift < min_shift) min_shift = segs[i].a_row_shift;
        if (segs[i].a_row_shift > max_shift) max_shift = segs[i].a_row_shift;
        if (segs[i].a_col_off + segs[i].klen > a_cols) a_cols = segs[i].a_col_off + segs[i].klen;
        if (segs[i].b_col_off + segs[i].klen > b_cols) b_cols = segs[i].b_col_off + segs[i].klen;
    }
    CSI_CHECK_ARG(a_cols <= lda && b_cols <= ldb, "segment exceeds the row pitch");
    const int BN = pick_bn(N);
    CUtensorMap tmA, tmB;
    const bf16* a_base = reinterpret_cast<const bf16*>(A) + (long long)min_shift * lda;
    int rc = make_map(&tmA, a_base, (long long)M + (max_shift - min_shift), a_cols, lda, TC_BM);
    if (rc) return rc;
    rc = make_map(&tmB, Bw, N, b_cols, ldb, BN);
    if (rc) return rc;
    NtParams p;
    p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.BN = BN;
    p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.drop_p = drop_p; p.drop_site = drop_site; p.rng = rng;
    p.row_base = -min_shift;
    uint32_t cols = 32;
    while ((int)cols < BN) cols <<= 1;
    p.tmem_cols = cols;
    const size_t smem = (size_t)TC_STAGES * (TC_BM * TC_BK * 2 + (size_t)BN * TC_BK * 2) + 1024;
    dim3 grid((M + TC_BM - 1) / TC_BM, (N + BN - 1) / BN);
    if (c_dtype == CSI_BF16) {
        CSI_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_nt_tc_kernel<bf16><<<grid, TC_THREADS, smem, ST(stream)>>>(tmA, tmB, p, sl);
    } else {
        CSI_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_nt_tc_kernel<float><<<grid, TC_THREADS, smem, ST(stream)>>>(tmA, tmB, p, sl);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// =================================================================================================================
//   csi_gemm_tn_tc : C[i*ldc + coff + q*cs] += sum_{m in chunk} A[m*lda + i] * Bv[(m+shift)*ldb + boff + q]
//                    Weight gradients (Linear and every Conv1d tap).  Both operands are read in their natural
//                    [tokens, channels] layout: the contraction runs over the row index, i.e. both UMMA operands are
//                    "MN-major" (a_major = b_major = 1) 128B-swizzled tiles filled by TMA boxes of 64 channels x 64 rows.
//                    The token range is split over blockIdx.z; partial tiles are reduced with fp32 atomics.
// =================================================================================================================
#define TN_STAGES 4
#define TN_BKM 64                 // token rows per pipeline stage (4 UMMA K-steps of 16)
#define TN_BOX_BYTES (TN_BKM * 128)

struct SegListTN { csi_seg_tn s[CSI_MAX_SEGS]; int n; };

// MN-major 128B-swizzled operand: 64 channels contiguous per 128 B row, rows = K; 8-row groups SBO = 1024 B apart;
// successive 64-channel chunks LBO = one TMA box (64 rows x 128 B) apart.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(TN_BOX_BYTES >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TnParams {
    float* C; int ldc; int cs; int M, Na, BN, chunk, qtiles;
    int row_base;
    uint32_t tmem_cols;
    csi_grp ig, qg;
};

__global__ void __launch_bounds__(TC_THREADS) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, TnParams p,
                                                                SegListTN segs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TN_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TN_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg_i = blockIdx.y / p.qtiles, qt = blockIdx.y % p.qtiles;
    const csi_seg_tn sg = segs.s[seg_i];
    const int i0 = blockIdx.x * TC_BM, q0 = qt * p.BN;
    if (q0 >= sg.nlen) return;                                     // uniform per CTA
    const int mbeg = blockIdx.z * p.chunk, mend = min(p.M, mbeg + p.chunk);
    const int nkb = (mend - mbeg + TN_BKM - 1) / TN_BKM;
    const int nbox_b = (p.BN + 63) / 64;
    const uint32_t a_bytes = 2 * TN_BOX_BYTES, b_bytes = (uint32_t)nbox_b * TN_BOX_BYTES;
    const uint32_t stage_bytes = a_bytes + b_bytes;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TN_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int stage = it % TN_STAGES;
                const uint32_t ph = (uint32_t)(it / TN_STAGES) & 1u;
                mbar_wait(&empty_bar[stage], ph ^ 1u);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                const int m = mbeg + it * TN_BKM;
                mbar_expect_tx(&full_bar[stage], stage_bytes);
                tma_load_2d(&tmA, &full_bar[stage], sa, i0, m);
                tma_load_2d(&tmA, &full_bar[stage], sa + TN_BOX_BYTES, i0 + 64, m);
                for (int bx = 0; bx < nbox_b; ++bx)
                    tma_load_2d(&tmB, &full_bar[stage], sa + a_bytes + (size_t)bx * TN_BOX_BYTES,
                                sg.b_col_off + q0 + bx * 64, m + sg.b_row_shift + p.row_base);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(TC_BM, p.BN) | (1u << 15) | (1u << 16);      // A and B MN-major
            for (int it = 0; it < nkb; ++it) {
                const int stage = it % TN_STAGES;
                const uint32_t ph = (uint32_t)(it / TN_STAGES) & 1u;
                mbar_wait(&full_bar[stage], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = make_mnmajor_desc(sa), bdesc = make_mnmajor_desc(sa + a_bytes);
                const int rows = min(TN_BKM, mend - (mbeg + it * TN_BKM));
                const int ksteps = (rows + 15) >> 4;           // rows past mend inside a 16-row step are real data of the
                for (int k = 0; k < ksteps; ++k)               // next chunk only if mend < M: chunk is a multiple of 16
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * (16 * 128 >> 4)), bdesc + (uint64_t)(k * (16 * 128 >> 4)),
                              idesc, (it | k) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int i = i0 + q * 32 + lane;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int ic = i < p.Na ? grp_to_compact(i, p.ig) : -1;
        float* crow = p.C + (size_t)(ic < 0 ? 0 : ic) * p.ldc + sg.c_off;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (ic >= 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int qq = q0 + c0 + j;
                    if (qq < sg.nlen) {
                        const int qc = grp_to_compact(qq, p.qg);
                        if (qc >= 0) atomicAdd(crow + (size_t)qc * p.cs, __uint_as_float(r[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

extern "C" int csi_gemm_tn_tc_ok(int lda, int ldb, int M, int Na, const csi_seg_tn* segs, int nseg) {
    if (M < 64 || Na < 1 || nseg < 1 || nseg > CSI_MAX_SEGS) return 0;
    if (lda % 8 || ldb % 8) return 0;
    for (int i = 0; i < nseg; ++i)
        if (segs[i].nlen <= 0 || segs[i].b_col_off % 8) return 0;
    return get_encode() != nullptr;
}

extern "C" int csi_gemm_tn_tc(const void* A, int lda, const void* Bv, int ldb, float* C, int ldc, int c_col_stride, int M,
                              int Na, const csi_seg_tn* segs, int nseg, csi_grp ig, csi_grp qg, void* stream) {
    CSI_CHECK_ARG(A && Bv && C && segs, "null pointer");
    CSI_CHECK_ARG(csi_gemm_tn_tc_ok(lda, ldb, M, Na, segs, nseg), "shape not eligible for the tcgen05 kernel");
    SegListTN sl;
    sl.n = nseg;
    int min_shift = 0, max_shift = 0, b_cols = 0, maxn = 0;
    for (int i = 0; i < nseg; ++i) {
        sl.s[i] = segs[i];
        if (segs[i].b_row_shift < min_shift) min_shift = segs[i].b_row_shift;
        if (segs[i].b_row_shift > max_shift) max_shift = segs[i].b_row_shift;
        if (segs[i].b_col_off + segs[i].nlen > b_cols) b_cols = segs[i].b_col_off + segs[i].nlen;
        if (segs[i].nlen > maxn) maxn = segs[i].nlen;
    }
    CSI_CHECK_ARG(b_cols <= ldb && Na <= lda, "segment exceeds the row pitch");
    // N tile: split the widest segment evenly into <= 256-wide tiles (multiples of 16)
    const int ntile = (maxn + 255) / 256;
    int BN = ((maxn + ntile - 1) / ntile + 15) & ~15;
    if (BN < 16) BN = 16;
    const int qtiles = (maxn + BN - 1) / BN;
    const int itiles = (Na + TC_BM - 1) / TC_BM;
    // token split: about two waves of CTAs, chunks a multiple of 64 rows and at least 512 rows
    const long long tiles = (long long)itiles * qtiles * nseg;
    int zs = (int)((2 * 148 + tiles - 1) / tiles);
    const int max_z = (M + 511) / 512;
    if (zs > max_z) zs = max_z;
    if (zs < 1) zs = 1;
    int chunk = ((M + zs - 1) / zs + TN_BKM - 1) / TN_BKM * TN_BKM;
    zs = (M + chunk - 1) / chunk;
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, A, M, Na, lda, TN_BKM);
    if (rc) return rc;
    const bf16* b_base = reinterpret_cast<const bf16*>(Bv) + (long long)min_shift * ldb;
    rc = make_map(&tmB, b_base, (long long)M + (max_shift - min_shift), b_cols, ldb, TN_BKM);
    if (rc) return rc;
    TnParams p;
    p.C = C; p.ldc = ldc; p.cs = c_col_stride; p.M = M; p.Na = Na; p.BN = BN; p.chunk = chunk; p.qtiles = qtiles;
    p.row_base = -min_shift;
    p.ig = ig; p.qg = qg;
    uint32_t cols = 32;
    while ((int)cols < BN) cols <<= 1;
    p.tmem_cols = cols;
    const int nbox_b = (BN + 63) / 64;
    const size_t smem = (size_t)TN_STAGES * (2 + nbox_b) * TN_BOX_BYTES + 1024;
    CSI_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(itiles, qtiles * nseg, zs);
    gemm_tn_tc_kernel<<<grid, TC_THREADS, smem, ST(stream)>>>(tmA, tmB, p, sl);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
