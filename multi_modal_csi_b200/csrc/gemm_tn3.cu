// tcgen05 weight-gradient GEMM with TAP SHARING (production path of csi_gemm_tn for bf16 operands).
//
//   C[i*ldc + coff_s + q*cs] += sum_m A[m*lda + i] * Bv[(m + shift_s)*ldb + boff_s + q]      for every segment s
//
// Both operands are read in their natural [tokens, channels] layout (contraction over the row index: MN-major UMMA
// operands, 128B swizzle, TMA boxes of 64 channels).  Segments that read the same columns of Bv at different row shifts
// (the taps of one Conv1d) are grouped: a CTA owns (128 channels of A) x (BN channels of Bv) x (up to 5 taps) and a
// chunk of the token range.  Per 64-token block it fetches the A tile once and ONE Bv tile with a halo of `span` rows;
// tap t's UMMA reads the Bv tile through a descriptor advanced by (shift_t - min_shift) rows of 128 B, and accumulates
// into its own TMEM slot.  So a k=5 conv issues 20 UMMAs per (A, Bv) stage instead of 4, and the operand traffic per
// FLOP drops 5x.  The partial sums of the token chunks go to a workspace registered for the stream (csi_gemm_tn_workspace) and
// are added to the reference [N, C, k] layout by tn3_reduce_kernel in a fixed order (bit-reproducible); without a workspace
// they are reduced with fp32 atomics from every CTA (round-1 path, kept for callers that own no scratch memory).
//
//   warp 0      TMA producer          warp 1      tcgen05.mma issuer (one elected lane, warp-uniform loop)
//   warps 2-9   epilogue: tcgen05.ld -> workspace slab (column-major, one 128-byte line per store) | red.global.add.f32
//               (two warps per TMEM lane quarter, alternate column groups)
#include "tc_common.cuh"
#include <stdlib.h>

#define ST(s) ((cudaStream_t)(s))
#define N3_THREADS 320
#define N3_STAGES 4
#define N3_BKM 64                 // token rows per pipeline stage (4 UMMA K-steps of 16)

struct Tn3Tap { int b_row_off; int c_off; };
struct Tn3Group { int b_col_off, nlen, min_shift, tap0, ntaps; };
struct Tn3Plan {
    Tn3Group g[CSI_MAX_SEGS];
    Tn3Tap t[CSI_MAX_SEGS];
    int ng;
};

struct Tn3Params {
    float* C; int ldc; int cs; int M, Na, BN, pitch, chunk, qtiles;
    int row_base, rows_b, nbox_b;
    uint32_t tmem_cols;
    csi_grp ig, qg;
    int dbg;
    // two-stage reduction (csi_gemm_tn_workspace): every CTA stores its raw accumulators as
    // ws[((tile * zs + z) * 128 + row) * ws_w + tmem column]; tn3_reduce_kernel sums the zs token chunks in a fixed order
    float* ws; int ws_w;
};

__device__ __forceinline__ bool elect_one_tn() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_tn(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_tn(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_tn(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_tn(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// MN-major 128B-swizzled operand: 64 channels contiguous per 128 B row, rows = K (tokens); 8-row groups SBO = 1024 B
// apart; successive 64-channel chunks LBO = one TMA box (box_bytes) apart.
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t box_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((box_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}


// Epilogue of one warp: TMEM lane quarter (warp % 4), alternate column passes with the other warp of the quarter.
// A pass covers NQ = 16*NLD consecutive q for all NT taps.  The lane (= row i) first writes its NT*NQ values to a
// per-warp shared-memory tile in the order [q][tap] -- for a Conv1d weight gradient in the reference [N, C, k] layout
// this is exactly the order in memory (index q*k + tap) -- then the warp walks the tile row by row so that consecutive
// lanes add to consecutive addresses: one red.global.add per 128-byte line instead of 32 scattered 4-byte atomics.
#define N3_TPITCH 81                                               // floats per staged row (odd: conflict-free)
template <int NT>
__device__ __forceinline__ void tn3_epilogue(const Tn3Params& p, const Tn3Plan& plan, int tap0, int g_nlen, int i0, int q0,
                                             uint32_t tmem_base, float* tbuf, int warp, int lane) {
    constexpr int NLD = NT == 1 ? 4 : (NT == 2 ? 2 : 1);
    constexpr int NQ = 16 * NLD, E = NQ * NT, NE = (E + 31) / 32;
    const int q = warp & 3, half = (warp - 2) >> 2;
    // row owned by this lane (for the TMEM read) and its compact index, broadcast later with shuffles
    const int irow = i0 + q * 32 + lane;
    const int my_ic = irow < p.Na ? grp_to_compact(irow, p.ig) : -1;
    const int npass = (p.BN + NQ - 1) / NQ;
    // dbg 8 / 16 (A/B): token chunks start at different column passes / rows, so that the CTAs of one output tile (which all
    // finish their mainloop together) do not queue on the same L2 lines
    const int prot = (p.dbg & 8) ? (int)(blockIdx.z % (unsigned)npass) : 0, rrot = (p.dbg & 16) ? (int)((blockIdx.z * 5u) & 31u) : 0;
    for (int pn = half; pn < npass; pn += 2) {
        int ps = pn + prot;
        if (ps >= npass) ps -= npass;
        const int c0 = ps * NQ;
        if (q0 + c0 >= g_nlen) continue;                           // warp-uniform
        uint32_t r[NT][NQ];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int l = 0; l < NLD; ++l) {
                if (c0 + l * 16 < p.BN) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.pitch + c0 + l * 16), &r[t][l * 16]);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[t][l * 16 + j] = 0u;
                }
            }
        // element e = j*NT + t of the staged row -> offset inside a row of C (fixed for the whole pass; -1 = masked)
        int off[NE];
#pragma unroll
        for (int u = 0; u < NE; ++u) {
            const int e = u * 32 + lane;
            const int j = e / NT, t = e - j * NT;
            const int qq = q0 + c0 + j;
            off[u] = -1;
            if (e < E && qq < g_nlen && c0 + j < p.BN) {
                const int qc = grp_to_compact(qq, p.qg);
                if (qc >= 0) off[u] = plan.t[tap0 + t].c_off + qc * p.cs;
            }
        }
        tmem_ld_wait();
        __syncwarp();                                              // previous pass has been read out of tbuf
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
            for (int t = 0; t < NT; ++t) tbuf[lane * N3_TPITCH + j * NT + t] = __uint_as_float(r[t][j]);
        __syncwarp();
        if (p.dbg & 1) continue;
#pragma unroll 4
        for (int r0 = 0; r0 < 32; ++r0) {
            const int rr = (r0 + rrot) & 31;
            const int ic = __shfl_sync(0xffffffffu, my_ic, rr);
            if (ic < 0) continue;                                  // warp-uniform
            float* crow = p.C + (size_t)ic * p.ldc;
#pragma unroll
            for (int u = 0; u < NE; ++u)
                if (off[u] >= 0) atomicAdd(crow + off[u], tbuf[rr * N3_TPITCH + u * 32 + lane]);
        }
    }
}

// Two-stage epilogue: the lane (= row i of the tile) copies its TMEM columns [0, ncols) to the workspace slab of this CTA,
// stored COLUMN-major (ws[slab][column][128 rows]): the 32 lanes of a warp hold 32 consecutive rows of one column, so every
// store instruction writes one whole 128-byte line.  The two warps of a lane quarter take alternate 32-column groups.
// No atomics: the token chunks are summed afterwards by tn3_reduce_kernel in a fixed order, so the weight gradient is
// bit-reproducible from run to run.
__device__ __forceinline__ void tn3_epilogue_ws(const Tn3Params& p, int ncols, int i0, uint32_t tmem_base, int warp, int lane) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int irow = i0 + q * 32 + lane;
    const size_t slab = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z;
    float* dst = p.ws + slab * (size_t)(TC_BM * p.ws_w) + (size_t)(q * 32 + lane);
    const bool ok = irow < p.Na;
    for (int c = half * 32; c < ncols; c += 64) {
        uint32_t r[32];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c + 16), r + 16);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(dst + (size_t)(c + j) * TC_BM, __uint_as_float(r[j]));
        }
    }
}

// Stage two: C[ic, c_off_t + qc * cs] += sum_z ws[tile, z, t * pitch + j, row].  One CTA = (tile, 32 rows, 8 workspace columns);
// warp w sums the token chunks z = w, w + 8, ... of those 8 columns (lane = row: every load instruction is one 128-byte line,
// 8-16 in flight per thread; the partial sums are L2-resident, written microseconds earlier), the eight per-warp sums are
// combined through shared memory in warp order, and thread (row, column) adds the result to C.  The order of every sum is
// fixed, so the gradient is bit-reproducible.  Grid = (row quarters, 8-column blocks, tiles).
__global__ void __launch_bounds__(256) tn3_reduce_kernel(Tn3Params p, const __grid_constant__ Tn3Plan plan, int zs, int itiles) {
    __shared__ float sm[8][8][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.z, ty = tile / itiles, it = tile - ty * itiles;        // slab order of the main kernel: (y, x, z)
    const int gi = ty / p.qtiles, qt = ty - gi * p.qtiles;
    const int q0 = qt * p.BN, g_nlen = plan.g[gi].nlen, tap0 = plan.g[gi].tap0, ntaps = plan.g[gi].ntaps;
    const int r0 = it * TC_BM + blockIdx.x * 32, c0 = blockIdx.y * 8;
    if (q0 >= g_nlen || r0 >= p.Na || c0 >= ntaps * p.pitch) return;                 // uniform per CTA
    {
        const int t0 = c0 / p.pitch, j0 = c0 - t0 * p.pitch;                         // 8 | pitch: the block lies inside one tap
        if (j0 >= p.BN || q0 + j0 >= g_nlen) return;
    }
    const size_t zstride = (size_t)TC_BM * p.ws_w;
    const float* src = p.ws + (size_t)tile * zs * zstride + (size_t)c0 * TC_BM + blockIdx.x * 32 + lane;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r0 + lane < p.Na) {
        int z = warp;
        for (; z + 8 < zs; z += 16) {
            float v0[8], v1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v0[u] = __ldcg(src + (size_t)z * zstride + u * TC_BM);
#pragma unroll
            for (int u = 0; u < 8; ++u) v1[u] = __ldcg(src + (size_t)(z + 8) * zstride + u * TC_BM);
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += v0[u];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += v1[u];
        }
        if (z < zs) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += __ldcg(src + (size_t)z * zstride + u * TC_BM);
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) sm[warp][u][lane] = a[u];
    __syncthreads();
    const int row = threadIdx.x >> 3, u = threadIdx.x & 7;
    const int c = c0 + u, t = c / p.pitch, j = c - t * p.pitch, qq = q0 + j, irow = r0 + row;
    if (irow >= p.Na || j >= p.BN || qq >= g_nlen) return;
    const int ic = grp_to_compact(irow, p.ig), qc = grp_to_compact(qq, p.qg);
    if (ic < 0 || qc < 0) return;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sm[w][u][row];
    // one add per address and launch (segments with disjoint outputs): the result does not depend on any ordering
    atomicAdd(p.C + (size_t)ic * p.ldc + plan.t[tap0 + t].c_off + qc * p.cs, v);
}

__global__ void __launch_bounds__(N3_THREADS, 1) gemm_tn_tc3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB, Tn3Params p,
                                                                    const __grid_constant__ Tn3Plan plan) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[N3_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[N3_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    if (p.dbg & 4) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.y / p.qtiles, qt = blockIdx.y - gi * p.qtiles;
    const int i0 = blockIdx.x * TC_BM, q0 = qt * p.BN;
    const int g_nlen = plan.g[gi].nlen, g_col = plan.g[gi].b_col_off, g_shift = plan.g[gi].min_shift;
    const int tap0 = plan.g[gi].tap0, ntaps = plan.g[gi].ntaps;
    if (q0 >= g_nlen) return;                                      // uniform per CTA
    const int mbeg = blockIdx.z * p.chunk, mend = min(p.M, mbeg + p.chunk);
    const int nkb = (mend - mbeg + N3_BKM - 1) / N3_BKM;
    const uint32_t a_box = N3_BKM * 128u, b_box = (uint32_t)p.rows_b * 128u;
    const uint32_t a_bytes = 2u * a_box, b_bytes = (uint32_t)p.nbox_b * b_box;
    const uint32_t stage_bytes = a_bytes + b_bytes;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < N3_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (p.dbg & 4) asm volatile("griddepcontrol.wait;" ::: "memory");      // PDL: the prologue above overlapped the previous kernel
    const uint32_t smem_u = smem_u32(smem), full_u = smem_u32(&full_bar[0]), empty_u = smem_u32(&empty_bar[0]);

    if (warp == 0) {
        const bool leader = elect_one_tn();
        int stage = 0;
        uint32_t ph = 1;
        for (int it = 0; it < nkb; ++it) {
            mbar_wait_tn(empty_u + 8u * stage, ph);
            if (leader) {
                const uint32_t sa = smem_u + (uint32_t)stage * stage_bytes, fb = full_u + 8u * stage;
                const int m = mbeg + it * N3_BKM;
                mbar_expect_tx_tn(fb, stage_bytes);
                tma_load_2d_tn(&tmA, fb, sa, i0, m);
                tma_load_2d_tn(&tmA, fb, sa + a_box, i0 + 64, m);
                for (int bx = 0; bx < p.nbox_b; ++bx)
                    tma_load_2d_tn(&tmB, fb, sa + a_bytes + (uint32_t)bx * b_box, g_col + q0 + bx * 64, m + g_shift + p.row_base);
            }
            if (++stage == N3_STAGES) { stage = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        const bool leader = elect_one_tn();
        const uint32_t idesc = make_idesc(TC_BM, p.BN) | (1u << 15) | (1u << 16);      // A and B MN-major
        int stage = 0;
        uint32_t ph = 0;
        for (int it = 0; it < nkb; ++it) {
            mbar_wait_tn(full_u + 8u * stage, ph);
            tc_fence_after();
            if (leader) {
                const uint32_t sa = smem_u + (uint32_t)stage * stage_bytes, sb = sa + a_bytes;
                const int rows = min(N3_BKM, mend - (mbeg + it * N3_BKM));
                const int ksteps = (rows + 15) >> 4;           // rows past mend inside a 16-row step: chunk is a multiple
                for (int t = 0; t < ntaps; ++t) {              // of 16 (real data of the next chunk never enters) or OOB zeros
                    const uint32_t tacc = tmem_base + (uint32_t)(t * p.pitch);
                    const uint32_t boff = (uint32_t)plan.t[tap0 + t].b_row_off * 128u;
                    const uint64_t adesc = make_mn_desc(sa, a_box), bdesc = make_mn_desc(sb + boff, b_box);
                    if (ksteps == 4) {
                        umma_bf16(tacc, adesc, bdesc, idesc, it ? 1u : 0u);
                        umma_bf16(tacc, adesc + 128, bdesc + 128, idesc, 1u);
                        umma_bf16(tacc, adesc + 256, bdesc + 256, idesc, 1u);
                        umma_bf16(tacc, adesc + 384, bdesc + 384, idesc, 1u);
                    } else {
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(tacc, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (it | k) ? 1u : 0u);
                    }
                }
                umma_commit_tn(empty_u + 8u * stage);
            }
            if (++stage == N3_STAGES) { stage = 0; ph ^= 1u; }
        }
        if (leader) umma_commit_tn(smem_u32(&tmem_full_bar));
    } else {
        mbar_wait(&tmem_full_bar, 0);                              // every UMMA has completed: the stage ring is free
        tc_fence_after();
        float* tbuf = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * (32 * N3_TPITCH);
        if (p.ws) { if (!(p.dbg & 2)) tn3_epilogue_ws(p, ntaps * p.pitch, i0, tmem_base, warp, lane); }
        else if (!(p.dbg & 2)) switch (ntaps) {
            case 1: tn3_epilogue<1>(p, plan, tap0, g_nlen, i0, q0, tmem_base, tbuf, warp, lane); break;
            case 2: tn3_epilogue<2>(p, plan, tap0, g_nlen, i0, q0, tmem_base, tbuf, warp, lane); break;
            case 3: tn3_epilogue<3>(p, plan, tap0, g_nlen, i0, q0, tmem_base, tbuf, warp, lane); break;
            case 4: tn3_epilogue<4>(p, plan, tap0, g_nlen, i0, q0, tmem_base, tbuf, warp, lane); break;
            default: tn3_epilogue<5>(p, plan, tap0, g_nlen, i0, q0, tmem_base, tbuf, warp, lane); break;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

static int g_num_sms_tn3 = 0;
static int g_tn_dbg = 0;
static int g_tn_pdl = -1;           // programmatic dependent launch: off by default (measured in-step: 4.79 ms with, 4.78 ms without), CSI_PDL=1 enables
extern "C" int csi_set_gemm_tn_pdl(int on) { g_tn_pdl = on ? 1 : 0; return CSI_OK; }
extern "C" int csi_set_tn_debug(int v) { g_tn_dbg = v; return CSI_OK; }

// ---- workspaces of the two-stage reduction, registered per (device, stream) by the caller that owns the memory
struct TnWorkspace { int dev; void* stream; float* ws; long long nfloats; };
static TnWorkspace g_tn_ws[32];
static int g_tn_nws = 0;
static std::mutex g_tn_ws_mu;
extern "C" int csi_gemm_tn_workspace(void* stream, float* ws, long long nfloats) {
    int dev = 0;
    CSI_CUDA(cudaGetDevice(&dev));
    CSI_CHECK_ARG(!ws || ((reinterpret_cast<uintptr_t>(ws) & 127) == 0 && nfloats > 0), "workspace must be 128-byte aligned");
    std::lock_guard<std::mutex> lk(g_tn_ws_mu);
    for (int i = 0; i < g_tn_nws; ++i)
        if (g_tn_ws[i].dev == dev && g_tn_ws[i].stream == stream) {
            if (ws) { g_tn_ws[i].ws = ws; g_tn_ws[i].nfloats = nfloats; }
            else g_tn_ws[i] = g_tn_ws[--g_tn_nws];
            return CSI_OK;
        }
    if (!ws) return CSI_OK;
    CSI_CHECK_ARG(g_tn_nws < 32, "too many registered streams");
    g_tn_ws[g_tn_nws++] = TnWorkspace{dev, stream, ws, nfloats};
    return CSI_OK;
}
static float* tn_workspace_for(void* stream, long long need) {
    if (g_tn_nws == 0) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_tn_ws_mu);
    for (int i = 0; i < g_tn_nws; ++i)
        if (g_tn_ws[i].dev == dev && g_tn_ws[i].stream == stream) return g_tn_ws[i].nfloats >= need ? g_tn_ws[i].ws : nullptr;
    return nullptr;
}

extern "C" int csi_gemm_tn_tc3(const void* A, int lda, const void* Bv, int ldb, float* C, int ldc, int c_col_stride, int M,
                               int Na, const csi_seg_tn* segs, int nseg, csi_grp ig, csi_grp qg, void* stream) {
    CSI_CHECK_ARG(A && Bv && C && segs, "null pointer");
    CSI_CHECK_ARG(nseg >= 1 && nseg <= CSI_MAX_SEGS && M >= 64 && lda % 8 == 0 && ldb % 8 == 0, "shape not eligible");
    // ---- tap groups: consecutive segments over the same Bv columns, row shifts within 16 rows, at most gmax taps
    int maxrun = 1, run = 1;
    for (int i = 1; i < nseg; ++i) {
        const bool same = segs[i].b_col_off == segs[i - 1].b_col_off && segs[i].nlen == segs[i - 1].nlen;
        run = same ? run + 1 : 1;
        if (run > maxrun) maxrun = run;
    }
    const int ngrp_per_run = (maxrun + 4) / 5;
    const int gmax = (maxrun + ngrp_per_run - 1) / ngrp_per_run;          // 1..5 taps per CTA
    Tn3Plan plan;
    plan.ng = 0;
    int min_shift = 0, max_shift = 0, b_cols = 0, maxn = 0, span = 0;
    for (int i = 0; i < nseg; ++i) {
        const csi_seg_tn s = segs[i];
        CSI_CHECK_ARG(s.nlen > 0 && s.b_col_off % 8 == 0, "bad segment");
        if (s.b_row_shift < min_shift) min_shift = s.b_row_shift;
        if (s.b_row_shift > max_shift) max_shift = s.b_row_shift;
        if (s.b_col_off + s.nlen > b_cols) b_cols = s.b_col_off + s.nlen;
        if (s.nlen > maxn) maxn = s.nlen;
        bool joined = false;
        if (plan.ng > 0) {
            Tn3Group& g = plan.g[plan.ng - 1];
            if (g.b_col_off == s.b_col_off && g.nlen == s.nlen && g.ntaps < gmax) {
                int lo = g.min_shift, hi = g.min_shift;
                for (int t = 0; t < g.ntaps; ++t) {
                    const int sh = g.min_shift + plan.t[g.tap0 + t].b_row_off;
                    if (sh > hi) hi = sh;
                }
                const int nlo = s.b_row_shift < lo ? s.b_row_shift : lo, nhi = s.b_row_shift > hi ? s.b_row_shift : hi;
                if (nhi - nlo <= 16) {
                    if (nlo != lo)
                        for (int t = 0; t < g.ntaps; ++t) plan.t[g.tap0 + t].b_row_off += lo - nlo;
                    g.min_shift = nlo;
                    plan.t[g.tap0 + g.ntaps].b_row_off = s.b_row_shift - nlo;
                    plan.t[g.tap0 + g.ntaps].c_off = s.c_off;
                    ++g.ntaps;
                    if (nhi - nlo > span) span = nhi - nlo;
                    joined = true;
                }
            }
        }
        if (!joined) {
            Tn3Group& g = plan.g[plan.ng];
            g.b_col_off = s.b_col_off; g.nlen = s.nlen; g.min_shift = s.b_row_shift; g.tap0 = i; g.ntaps = 1;
            plan.t[i].b_row_off = 0; plan.t[i].c_off = s.c_off;
            ++plan.ng;
        }
    }
    CSI_CHECK_ARG(b_cols <= ldb && Na <= lda, "segment exceeds the row pitch");
    int gtaps = 1;
    for (int g = 0; g < plan.ng; ++g) if (plan.g[g].ntaps > gtaps) gtaps = plan.g[g].ntaps;
    // N tile: every tap of a group owns a TMEM slot of `pitch` columns (multiple of 32), gtaps * pitch <= 512
    int bn_cap = (512 / gtaps) & ~31;
    if (bn_cap > 256) bn_cap = 256;
    const int ntile = (maxn + bn_cap - 1) / bn_cap;
    int BN = ((maxn + ntile - 1) / ntile + 15) & ~15;
    if (BN < 16) BN = 16;
    const int pitch = (BN + 31) & ~31;
    const int qtiles = (maxn + BN - 1) / BN;
    const int itiles = (Na + TC_BM - 1) / TC_BM;
    if (g_num_sms_tn3 == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_num_sms_tn3, cudaDevAttrMultiProcessorCount, dev));
    }
    // token split: one wave of CTAs (one CTA per SM: the accumulators fill TMEM), chunks a multiple of 64 rows
    const long long tiles = (long long)itiles * qtiles * plan.ng;
    int zs = (int)(g_num_sms_tn3 / tiles);
    const int max_z = (M + 511) / 512;
    if (zs > max_z) zs = max_z;
    if (zs < 1) zs = 1;
    int chunk = ((M + zs - 1) / zs + N3_BKM - 1) / N3_BKM * N3_BKM;
    zs = (M + chunk - 1) / chunk;
    const int rows_b = N3_BKM + ((span + 7) & ~7);
    const int nbox_b = (BN + 63) / 64;
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, A, M, Na, lda, N3_BKM);
    if (rc) return rc;
    const bf16* b_base = reinterpret_cast<const bf16*>(Bv) + (long long)min_shift * ldb;
    rc = make_map(&tmB, b_base, (long long)M + (max_shift - min_shift), b_cols, ldb, rows_b);
    if (rc) return rc;
    Tn3Params p;
    p.C = C; p.ldc = ldc; p.cs = c_col_stride; p.M = M; p.Na = Na; p.BN = BN; p.pitch = pitch; p.chunk = chunk; p.qtiles = qtiles;
    p.row_base = -min_shift; p.rows_b = rows_b; p.nbox_b = nbox_b;
    if (g_tn_pdl < 0) { const char* e = getenv("CSI_PDL"); g_tn_pdl = (e && e[0] == '1') ? 1 : 0; }
    p.ig = ig; p.qg = qg; p.dbg = g_tn_dbg | (g_tn_pdl ? 4 : 0);
    uint32_t cols = 32;
    while ((int)cols < gtaps * pitch) cols <<= 1;
    p.tmem_cols = cols;
    p.ws_w = gtaps * pitch;
    p.ws = tn_workspace_for(stream, (long long)tiles * zs * TC_BM * p.ws_w);
    const size_t smem = (size_t)N3_STAGES * (2 * N3_BKM * 128 + (size_t)nbox_b * rows_b * 128) + 1024;
    CSI_CHECK_ARG(smem <= 227 * 1024, "stage does not fit in shared memory");
    CSI_CUDA(cudaFuncSetAttribute(gemm_tn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(itiles, qtiles * plan.ng, zs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(N3_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ST(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_tn_pdl;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (!(g_tn_dbg & 64)) CSI_CUDA(cudaLaunchKernelEx(&cfg, gemm_tn_tc3_kernel, tmA, tmB, p, plan));
    CSI_LAUNCH_CHECK();
    if (p.ws && !(g_tn_dbg & 32)) {
        // stage two (plain stream order: it starts once every partial sum is in memory)
        tn3_reduce_kernel<<<dim3(TC_BM / 32, p.ws_w / 8, (unsigned)tiles), 256, 0, ST(stream)>>>(p, plan, zs, itiles);
        CSI_LAUNCH_CHECK();
    }
    return CSI_OK;
}
