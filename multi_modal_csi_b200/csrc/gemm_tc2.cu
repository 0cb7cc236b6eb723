// Persistent tcgen05 NT GEMM (the production path of csi_gemm_nt for bf16 operands).
//
//   C[m,n] = sum_seg A[(m+shift)*lda + aoff + q] * B[n*ldb + boff + q]   (+bias) -> dropout -> (+residual)
//
// One CTA per SM walks 128 x BN output tiles (n fastest, so the CTAs that run together share an A tile in L2).
//   warp 0      TMA producer: K-major 128B-swizzled A/B stages, ring of NS stages, mbarrier tx-count
//   warp 1      tcgen05.mma issuer (UMMA 128 x BN x 16, bf16 -> fp32), two TMEM accumulators so the next tile's
//               main loop runs while the previous tile is drained
//   warps 2-5   epilogue: tcgen05.ld (one TMEM lane quarter per warp) -> bias / Philox dropout / fp32 residual
//               (prefetched one panel ahead) -> 128B-swizzled smem panel -> TMA store (clips the M/N tails)
#include "tc_common.cuh"

#define ST(s) ((cudaStream_t)(s))
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
#define T2_THREADS 192
#define T2_MAX_STAGES 8

struct SegList2 { csi_seg s[CSI_MAX_SEGS]; int n; };

struct Nt2Params {
    int M, N, BN, ntn, ntiles, nstages;          // ntiles = cluster work items (super m-tiles x n-tiles)
    void* C; int ldc;
    const float* bias; const float* residual; int ldr;
    float drop_p; unsigned drop_site; const unsigned long long* rng;
    int row_base;
};

// PW = columns per epilogue panel (one 128-byte smem row): 64 for bf16 output, 32 for fp32 output
template <typename TC, bool RES, int CL>
__global__ void __launch_bounds__(T2_THREADS, 1) gemm_nt_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmC, Nt2Params p,
                                                                    SegList2 segs) {
    constexpr int PW = 128 / (int)sizeof(TC);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[T2_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[T2_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float sbias[2][256];

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = TC_BM * TC_BK * 2, b_bytes = (uint32_t)p.BN * TC_BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int NS = p.nstages;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const uint32_t bslice = b_bytes / CL;
    uint8_t* stage_base = smem;
    uint8_t* cstage = smem + (size_t)NS * stage_bytes;            // 4 warps x 2 buffers x (32 rows x 128 B)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                          // peers' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = cid; tile < p.ntiles; tile += ncl) {
                const int m0 = ((tile / p.ntn) * CL + (int)crank) * TC_BM, n0 = (tile % p.ntn) * p.BN;
                for (int s = 0; s < segs.n; ++s) {
                    const csi_seg sg = segs.s[s];
                    for (int k0 = 0; k0 < sg.klen; k0 += TC_BK, ++it) {
                        const int stage = it % NS;
                        const uint32_t ph = (uint32_t)(it / NS) & 1u;
                        mbar_wait(&empty_bar[stage], ph ^ 1u);
                        uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full_bar[stage], stage_bytes);
                        tma_load_2d(&tmA, &full_bar[stage], sa, sg.a_col_off + k0, m0 + sg.a_row_shift + p.row_base);
                        if (CL > 1)
                            tma_load_2d_mc(&tmB, &full_bar[stage], sa + a_bytes + crank * bslice, sg.b_col_off + k0,
                                           n0 + (int)crank * (p.BN / CL), cmask);
                        else
                            tma_load_2d(&tmB, &full_bar[stage], sa + a_bytes, sg.b_col_off + k0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(TC_BM, p.BN);
            int it = 0, ti = 0;
            for (int tile = cid; tile < p.ntiles; tile += ncl, ++ti) {
                const int acc = ti & 1;
                const uint32_t aph = (uint32_t)(ti >> 1) & 1u;
                mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);         // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * 256);
                uint32_t first = 1;
                for (int s = 0; s < segs.n; ++s) {
                    const csi_seg sg = segs.s[s];
                    for (int k0 = 0; k0 < sg.klen; k0 += TC_BK, ++it) {
                        const int stage = it % NS;
                        const uint32_t ph = (uint32_t)(it / NS) & 1u;
                        mbar_wait(&full_bar[stage], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * stage_bytes);
                        const uint64_t adesc = make_kmajor_desc(sa), bdesc = make_kmajor_desc(sa + a_bytes);
                        const int ksteps = min(TC_BK, sg.klen - k0) >> 4;
                        for (int k = 0; k < ksteps; ++k) {
                            umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                        if (CL > 1) umma_commit_mc(&empty_bar[stage], cmask);   // frees this stage in every CTA of the cluster
                        else umma_commit(&empty_bar[stage]);
                    }
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3, ew = warp - 2;
        const int et = threadIdx.x - 64;                            // 0..127 among the epilogue threads
        const bool drop = p.drop_p > 0.f;
        DropCtx dc;
        if (drop) dc = drop_ctx(p.rng, p.drop_p);
        const int ld8 = ((p.N + 15) & ~15) >> 3;
        uint8_t* mybuf = cstage + (size_t)ew * 2 * 4096;
        int ti = 0, sbuf = 0;
        for (int tile = cid; tile < p.ntiles; tile += ncl, ++ti) {
            const int acc = ti & 1;
            const uint32_t aph = (uint32_t)(ti >> 1) & 1u;
            const int m0 = ((tile / p.ntn) * CL + (int)crank) * TC_BM, n0 = (tile % p.ntn) * p.BN;
            const int m = m0 + q * 32 + lane;
            if (p.bias) {
                for (int i = et; i < p.BN; i += 128) sbias[acc][i] = (n0 + i) < p.N ? p.bias[n0 + i] : 0.f;
                named_bar_sync(1, 128);
            }
            const float* rrow = (RES && m < p.M) ? p.residual + (size_t)m * p.ldr : nullptr;
            float rnext[RES ? PW : 1];
            auto fetch_res = [&](int pc0) {
                if constexpr (!RES) return;
#pragma unroll
                for (int j = 0; j < (RES ? PW : 0); j += 4) {
                    const int n = n0 + pc0 + j;
                    if (rrow && n + 3 < p.N) {
                        const float4 v = *reinterpret_cast<const float4*>(rrow + n);
                        rnext[j] = v.x; rnext[j + 1] = v.y; rnext[j + 2] = v.z; rnext[j + 3] = v.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u) rnext[j + u] = (rrow && n + u < p.N) ? rrow[n + u] : 0.f;
                    }
                }
            };
            if constexpr (RES) fetch_res(0);
            mbar_wait(&tmem_full_bar[acc], aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
            for (int pc0 = 0; pc0 < p.BN; pc0 += PW) {
                float v[PW];
#pragma unroll
                for (int c = 0; c < PW; c += 16) {
                    uint32_t r[16];
                    if (pc0 + c < p.BN) tmem_ld16(tacc + (uint32_t)(pc0 + c), r);
                    else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[j] = 0u;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c + j] = __uint_as_float(r[j]);
                }
                float res[RES ? PW : 1];
                if constexpr (RES) {
#pragma unroll
                    for (int j = 0; j < PW; ++j) res[j] = rnext[j];
                    if (pc0 + PW < p.BN) fetch_res(pc0 + PW);
                }
#pragma unroll
                for (int g8 = 0; g8 < PW / 8; ++g8) {
                    float ks[8];
                    if (drop) drop_scales8(dc, p.drop_site, (unsigned long long)m * ld8 + ((n0 + pc0) >> 3) + g8, ks);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float x = v[g8 * 8 + j];
                        if (p.bias) x += sbias[acc][min(pc0 + g8 * 8 + j, 255)];
                        if (drop) x *= ks[j];
                        if constexpr (RES) x += res[g8 * 8 + j];
                        v[g8 * 8 + j] = x;
                    }
                }
                if (pc0 + PW <= p.BN) {
                    // full panel: stage the 32 x 128 B slice of this warp (row = lane) with the TMA 128B swizzle and
                    // bulk-store it (the tensor map clips rows >= M and columns >= N)
                    if (lane == 0) tma_store_wait_read<1>();        // the buffer used two panels ago has been read
                    __syncwarp();
                    uint8_t* buf = mybuf + (size_t)sbuf * 4096;
                    uint8_t* rowp = buf + lane * 128;
#pragma unroll
                    for (int c16 = 0; c16 < 8; ++c16) {
                        uint4 u;
                        if constexpr (sizeof(TC) == 2) {
                            u.x = pack_bf16x2(v[c16 * 8 + 0], v[c16 * 8 + 1]); u.y = pack_bf16x2(v[c16 * 8 + 2], v[c16 * 8 + 3]);
                            u.z = pack_bf16x2(v[c16 * 8 + 4], v[c16 * 8 + 5]); u.w = pack_bf16x2(v[c16 * 8 + 6], v[c16 * 8 + 7]);
                        } else {
                            u.x = __float_as_uint(v[c16 * 4 + 0]); u.y = __float_as_uint(v[c16 * 4 + 1]);
                            u.z = __float_as_uint(v[c16 * 4 + 2]); u.w = __float_as_uint(v[c16 * 4 + 3]);
                        }
                        *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) = u;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, buf, n0 + pc0, m0 + q * 32);
                        tma_store_commit();
                    }
                    sbuf ^= 1;
                } else if (m < p.M) {
                    // ragged last panel of the tile (BN is a multiple of 16, not of the panel width): direct stores
                    TC* crow = reinterpret_cast<TC*>(p.C) + (size_t)m * p.ldc;
#pragma unroll
                    for (int j = 0; j < PW; j += 2) {
                        const int n = n0 + pc0 + j;
                        if (pc0 + j < p.BN && n < p.N) {
                            if (n + 1 < p.N) st2<TC>(crow + n, make_float2(v[j], v[j + 1]));
                            else stf<TC>(crow + n, v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                          // no CTA leaves while a peer may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int pick_bn2(int N) {
    const int tiles = (N + 255) / 256;
    int bn = ((N + tiles - 1) / tiles + 15) & ~15;      // multiple of 16: each of the 2 multicast slices is whole 8-row swizzle atoms
    if (bn < 16) bn = 16;
    return bn;
}

static int g_num_sms = 0;
static int g_cluster2 = 0;     // measured: 2-CTA multicast does not cut L2 traffic on B200 (dedup window), lockstep costs 8%
extern "C" int csi_set_gemm_cluster(int on) { g_cluster2 = on ? 1 : 0; return CSI_OK; }

extern "C" int csi_gemm_nt_tc2(const void* A, int lda, const void* Bw, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                               const csi_seg* segs, int nseg, const float* bias, const float* residual, int ldr,
                               float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(A && Bw && C && segs, "null pointer");
    CSI_CHECK_ARG(!(drop_p > 0.f) || rng, "dropout needs rng");
    const int es = c_dtype == CSI_BF16 ? 2 : 4;
    CSI_CHECK_ARG((ldc * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "C must be 16-byte aligned with a 16-byte row pitch");
    CSI_CHECK_ARG(!residual || (ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0), "residual must be 16-byte aligned");
    SegList2 sl;
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
    if (g_num_sms == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int BN = pick_bn2(N);
    CUtensorMap tmA, tmB, tmC;
    const bf16* a_base = reinterpret_cast<const bf16*>(A) + (long long)min_shift * lda;
    int rc = make_map(&tmA, a_base, (long long)M + (max_shift - min_shift), a_cols, lda, TC_BM);
    if (rc) return rc;
    const int mtiles = (M + TC_BM - 1) / TC_BM;
    const int CL = (g_cluster2 && mtiles >= 4) ? 2 : 1;     // CTA pairs share the weight tile through TMA multicast
    rc = make_map(&tmB, Bw, N, b_cols, ldb, BN / CL);
    if (rc) return rc;
    rc = make_map_ex(&tmC, C, M, N, ldc, 32, 128 / es, es);
    if (rc) return rc;
    Nt2Params p;
    p.M = M; p.N = N; p.BN = BN; p.C = C; p.ldc = ldc;
    p.ntn = (N + BN - 1) / BN;
    p.ntiles = p.ntn * ((mtiles + CL - 1) / CL);
    p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.drop_p = drop_p; p.drop_site = drop_site; p.rng = rng;
    p.row_base = -min_shift;
    const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)BN * TC_BK * 2;
    const size_t fixed = 1024 + 4 * 2 * 4096;
    int ns = (int)((220 * 1024 - fixed) / stage_bytes);
    if (ns > T2_MAX_STAGES) ns = T2_MAX_STAGES;
    if (ns < 2) ns = 2;
    p.nstages = ns;
    const size_t smem = fixed + (size_t)ns * stage_bytes;
    int grid = (g_num_sms / CL) * CL;
    if (p.ntiles * CL < grid) grid = p.ntiles * CL;
#define LAUNCH(TC, RES, CLV)                                                                                            \
    do {                                                                                                                \
        CSI_CUDA(cudaFuncSetAttribute(gemm_nt_tc2_kernel<TC, RES, CLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cudaLaunchConfig_t cfg = {};                                                                                    \
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ST(stream); \
        cudaLaunchAttribute at[1];                                                                                      \
        at[0].id = cudaLaunchAttributeClusterDimension;                                                                 \
        at[0].val.clusterDim.x = CLV; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                           \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                               \
        CSI_CUDA(cudaLaunchKernelEx(&cfg, gemm_nt_tc2_kernel<TC, RES, CLV>, tmA, tmB, tmC, p, sl));                     \
    } while (0)
    CSI_CHECK_ARG(!(residual && c_dtype == CSI_BF16), "residual is only fused for fp32 output");
    if (CL == 2) {
        if (c_dtype == CSI_BF16) LAUNCH(bf16, false, 2);
        else if (residual) LAUNCH(float, true, 2);
        else LAUNCH(float, false, 2);
    } else {
        if (c_dtype == CSI_BF16) LAUNCH(bf16, false, 1);
        else if (residual) LAUNCH(float, true, 1);
        else LAUNCH(float, false, 1);
    }
#undef LAUNCH
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
