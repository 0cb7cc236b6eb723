// Persistent tcgen05 NT GEMM with TAP SHARING (production path of csi_gemm_nt for bf16 operands).
//
//   C[m,n] = sum_seg A[(m+shift_seg)*lda + aoff_seg + q] * B[n*ldb + boff_seg + q]   (+bias) -> dropout -> (+residual)
//
// The GEMMs of this model are bound by L2 -> SM operand traffic, not by the tensor pipe (N is only 128..270, so a
// 128 x BN tile re-reads its A tile for every tap of a Conv1d).  Segments that read the SAME columns of A at different
// row shifts (the taps of one Conv1d / of one branch in the data-gradient) are therefore grouped: per 64-channel block
// the A tile is fetched ONCE as a (128 + span) x 64 box and every tap's UMMA reads it through a shared-memory descriptor
// whose start address is advanced by (shift - min_shift) rows of 128 B.  The 128B swizzle is a function of the absolute
// shared-memory address (TMA writes and UMMA reads agree), so a row-shifted view of the same stage is a valid K-major
// operand.  B (weights) tiles stream through their own ring, one per (tap, channel block).
//
// Banded calls (csi_gemm_nt_banded): every segment carries the band of output columns it contributes to (its weights are zero
// elsewhere), column tiles are cut at the band boundaries (T3Params::tab_n0 / tab_bn) and the producer and issue loops skip the
// taps whose band does not meet the tile -- the three Conv1d branches of an encoder run as one GEMM with N = 3*Dp.
//
//   warp 0      TMA producer: A ring (box rows = 128 + span) and B ring (BN x 64), mbarrier tx-count
//   warp 1      tcgen05.mma issuer (UMMA 128 x BN x 16, bf16 -> fp32), two TMEM accumulators
//   warps 2-5   epilogue: tcgen05.ld -> bias / Philox dropout / fp32 residual -> swizzled smem panel -> TMA store
#include "tc_common.cuh"
#include <stdlib.h>

#define ST(s) ((cudaStream_t)(s))
#define T3_THREADS 320
#define T3_EPI_THREADS 256
#define T3_MAX_STAGES 16
#define T3_MAX_NTAB 12

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_u(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster run ONE 256 x BN tile; each stages its own 128 rows of A
// and HALF of the weight tile, the leader (cluster rank 0) issues every UMMA for both tensor cores
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope): a .cluster-scope release costs a cluster-wide memory fence per arrive (measured:
    // +3000 clk per tile in the epilogue); the TMEM reads it orders are already fenced by tcgen05.fence::before_thread_sync
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the completion bytes of a pair load are signalled on the LEADER's barrier (bar = shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

struct T3Tap { short a_row_off; short pad; int b_col_off; int n_lo, n_hi; };       // [n_lo, n_hi): output columns the tap contributes to
struct T3Group { int a_col_off, klen, min_shift, tap0, ntaps; };
struct T3Plan {
    T3Group g[CSI_MAX_SEGS];
    T3Tap t[CSI_MAX_SEGS];
    int ng;
};

struct Nt3Params {
    int M, N, BN, ntn, ntiles, nsa, nsb, a_rows;
    void* C; int ldc;
    const float* bias; const float* residual; int ldr;
    float drop_p; unsigned drop_site; const unsigned long long* rng;
    int row_base, desc_mode, pdl;
    // tail-wave split: tiles [0, nfull) are 128 x BN; the M tiles of the last, partly filled wave are cut into `npiece`
    // column pieces of BN2 (= one or two store panels) so that every SM gets a short piece instead of a few SMs a whole tile
    int nfull, npiece, BN2, mt_tail;
    int epi_bufs;                     // staging buffers per epilogue warp (1 or 2): one leaves room for a deeper B ring
    int ksub;                         // 64-channel blocks per pipeline stage (1 or 2), see the mainloop comment
    // weight-stationary mode (short-K calls): the B ring holds EVERY weight stage of a tile (nsb = stages per tile); a CTA
    // (pair) keeps one column tile for its whole life, loads the weights once and then only streams A.  Without it a
    // K = 272 call re-fetches 130 KB of weights per 70 KB of activations and is bound by the SM's share of L2 bandwidth.
    int resident, mtiles;
    long long* dbg;                   // optional phase clocks of one mid-grid CTA (scripts/gemm_phases.py), NULL in production
    // banded calls (csi_gemm_nt_banded: the three Conv1d branches of an encoder as ONE GEMM): column tiles follow the band
    // boundaries, so their widths differ (272 = 144 + 128); ntab > 0 replaces the uniform (tile % ntn) * BN rule
    int ntab;
    short tab_n0[T3_MAX_NTAB], tab_bn[T3_MAX_NTAB];
};
// The probes sit in the single-threaded issue loops, whose latency is the kernel's critical path: they are compiled in only
// with -DCSI_T3_DEBUG (make EXTRA=-DCSI_T3_DEBUG).
#ifdef CSI_T3_DEBUG
#define T3_CLK(slot) do { if (dbg_on && (slot) < 4096) p.dbg[(slot)] = clock64(); } while (0)
#define T3_DBG_ON(cond) (p.dbg != nullptr && (cond))
#else
#define T3_CLK(slot) do { } while (0)
#define T3_DBG_ON(cond) false
#endif

struct T3Tile { int mt, n0, bn; };
__device__ __forceinline__ T3Tile t3_decode(const Nt3Params& p, int tile, int n_res) {
    T3Tile t;
    if (p.ntab) {
        t.mt = tile / p.ntab;
        const int nt = tile - t.mt * p.ntab;
        t.n0 = p.tab_n0[nt]; t.bn = p.tab_bn[nt];
    } else if (p.resident) {
        t.mt = tile; t.n0 = n_res * p.BN; t.bn = p.BN;
    } else if (tile < p.nfull) {
        t.mt = tile / p.ntn;
        t.n0 = (tile - t.mt * p.ntn) * p.BN;
        t.bn = p.BN;
    } else {
        const int q = tile - p.nfull, m = q / p.npiece;
        t.mt = p.mt_tail + m;
        t.n0 = (q - m * p.npiece) * p.BN2;
        t.bn = min(p.BN2, p.BN - t.n0);                  // the split is only used with one column tile per row tile
    }
    return t;
}

// K-major SW128 descriptor with an explicit matrix-base-offset field (bits [49,52))
__device__ __forceinline__ uint64_t make_kmajor_desc_bo(uint32_t saddr, uint32_t base_off) {
    return make_kmajor_desc(saddr) | ((uint64_t)(base_off & 7u) << 49);
}

template <typename TC, bool RES, bool PAIR>
__global__ void __launch_bounds__(T3_THREADS, 1) gemm_nt_tc3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmB2,
                                                                    const __grid_constant__ CUtensorMap tmC,
                                                                    const __grid_constant__ CUtensorMap tmCt,
                                                                    const __grid_constant__ CUtensorMap tmR,
                                                                    const __grid_constant__ CUtensorMap tmRt, Nt3Params p,
                                                                    const __grid_constant__ T3Plan plan) {
    constexpr int PW = 128 / (int)sizeof(TC);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[T3_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_a[T3_MAX_STAGES];
    __shared__ __align__(8) uint64_t full_b[T3_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_b[T3_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float sbias[2][256];
    __shared__ __align__(8) uint64_t res_full[T3_EPI_THREADS / 32][2];   // RES: residual panel ring of every epilogue warp

    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel may start its prologue
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // PAIR: tiles are 256 x BN per cluster of two CTAs; `rank` 0 is the leader.  The tile loop runs over cluster indices.
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // weight-stationary: unit u owns column tile u % ntn and walks the row tiles u / ntn, u / ntn + (units with that column tile), ...
    const int n_res = p.resident ? unit % p.ntn : 0;
    const int tile_first = p.resident ? unit / p.ntn : unit;
    const int tile_step = p.resident ? (nunits - n_res + p.ntn - 1) / p.ntn : nunits;
    const int tile_end = p.resident ? p.mtiles : p.ntiles;
    const bool resident = p.resident != 0;
    constexpr int TILE_M = PAIR ? 2 * TC_BM : TC_BM;
    // A stage of the rings holds KSUB consecutive 64-channel blocks (each its own TMA box / swizzle atom column): with
    // narrow tiles (BN <= 160) one block is only 4 UMMAs of <= 80 clk, less than the ~500 clk the single-threaded issue
    // loop needs per barrier round trip, so two blocks share one round trip.
    const uint32_t a_bytes = (uint32_t)p.a_rows * 128u, b_bytes = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * 128u;   // per CTA
    const int NSA = p.nsa, NSB = p.nsb, KSUB = p.ksub;
    const uint32_t a_sbytes = a_bytes * (uint32_t)KSUB, b_sbytes = b_bytes * (uint32_t)KSUB;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)NSA * a_sbytes;
    uint8_t* cstage = b_ring + (size_t)NSB * b_sbytes;             // 8 warps x 2 buffers x (32 rows x 128 B)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmCt);
        for (int s = 0; s < NSA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < NSB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        // PAIR: the epilogue warps of both CTAs release an accumulator on the leader's barrier
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], (PAIR ? 2 : 1) * T3_EPI_THREADS / 32); }
        if constexpr (RES)
            for (int w = 0; w < T3_EPI_THREADS / 32; ++w) { mbar_init(&res_full[w][0], 1); mbar_init(&res_full[w][1], 1); }
        fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_pair(&tmem_base_smem, 512); else tmem_alloc(&tmem_base_smem, 512); }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();       // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; no
    // global memory is touched before the previous grid has completed and flushed
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

    // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the barriers) and issue the
    // asynchronous instructions from one elected lane: the issue path is a single instruction stream whose latency
    // per stage must stay below the stage's tensor time (4 x 72 clk for a 128x144 tile), so it carries no integer
    // division (stage/phase are running counters) and no per-lane divergence.
    if (warp == 0) {
        int sa = 0, sb = 0;
        uint32_t pa = 1, pb = 1;                                 // parity to wait for on the empty barriers
        const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
        // PAIR: both CTAs load into their own rings, every completion is counted on the LEADER's full barriers, and only
        // the leader posts the expected byte count (of both CTAs)
        const uint32_t full_a_u = PAIR ? mapa_u32(smem_u32(&full_a[0]), 0u) : smem_u32(&full_a[0]), empty_a_u = smem_u32(&empty_a[0]);
        const uint32_t full_b_u = PAIR ? mapa_u32(smem_u32(&full_b[0]), 0u) : smem_u32(&full_b[0]), empty_b_u = smem_u32(&empty_b[0]);
        const uint32_t full_a_l = smem_u32(&full_a[0]), full_b_l = smem_u32(&full_b[0]);
        const bool leader = elect_one();
        const bool post = leader && rank == 0;
        constexpr uint32_t NCTA = PAIR ? 2u : 1u;
        const bool dbg_on = T3_DBG_ON(blockIdx.x == (gridDim.x / 2 & ~1u) && leader); (void)dbg_on;
        int dti = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++dti) {
            T3_CLK(2000 + 8 * dti);
            const T3Tile tt = t3_decode(p, tile, n_res);
            const int m0 = tt.mt * TILE_M + (int)rank * TC_BM + p.row_base, n0 = tt.n0 + (PAIR ? (int)rank * (tt.bn >> 1) : 0);
            const bool piece = tile >= p.nfull;                  // pieces fetch BN2-row weight boxes through their own map
            const CUtensorMap* tmb = piece ? &tmB2 : &tmB;
            const uint32_t bb = piece ? (uint32_t)(PAIR ? p.BN2 / 2 : p.BN2) * 128u : b_bytes;
            for (int gi = 0; gi < plan.ng; ++gi) {
                const int g_col = plan.g[gi].a_col_off, g_klen = plan.g[gi].klen, g_row = m0 + plan.g[gi].min_shift;
                const int tap0 = plan.g[gi].tap0, ntaps = plan.g[gi].ntaps;
                for (int k0 = 0; k0 < g_klen; k0 += KSUB * TC_BK) {
                    const int nsub = (KSUB == 2 && g_klen - k0 > TC_BK) ? 2 : 1;
                    mbar_wait_u(empty_a_u + 8u * sa, pa);
                    if (post) mbar_expect_tx_u(full_a_l + 8u * sa, a_bytes * (uint32_t)nsub * NCTA);
                    if (leader) {
                        if constexpr (PAIR) {
                            tma_load_2d_pair(&tmA, full_a_u + 8u * sa, a_ring_u + (uint32_t)sa * a_sbytes, g_col + k0, g_row);
                            if (nsub == 2)
                                tma_load_2d_pair(&tmA, full_a_u + 8u * sa, a_ring_u + (uint32_t)sa * a_sbytes + a_bytes, g_col + k0 + TC_BK, g_row);
                        } else {
                            tma_load_2d_u(&tmA, full_a_u + 8u * sa, a_ring_u + (uint32_t)sa * a_sbytes, g_col + k0, g_row);
                            if (nsub == 2)
                                tma_load_2d_u(&tmA, full_a_u + 8u * sa, a_ring_u + (uint32_t)sa * a_sbytes + a_bytes, g_col + k0 + TC_BK, g_row);
                        }
                    }
                    if (++sa == NSA) { sa = 0; pa ^= 1u; }
                    for (int t = 0; t < ntaps; ++t) {
                        if (tt.n0 >= plan.t[tap0 + t].n_hi || tt.n0 + tt.bn <= plan.t[tap0 + t].n_lo) continue;   // band: tap absent from these columns
                        if (resident) {                          // weights are loaded with the first tile only
                            if (tile != tile_first) continue;
                        } else mbar_wait_u(empty_b_u + 8u * sb, pb);
                        if (post) mbar_expect_tx_u(full_b_l + 8u * sb, bb * (uint32_t)nsub * NCTA);
                        if (leader) {
                            const int bcol = plan.t[tap0 + t].b_col_off + k0;
                            if constexpr (PAIR) {
                                tma_load_2d_pair(tmb, full_b_u + 8u * sb, b_ring_u + (uint32_t)sb * b_sbytes, bcol, n0);
                                if (nsub == 2)
                                    tma_load_2d_pair(tmb, full_b_u + 8u * sb, b_ring_u + (uint32_t)sb * b_sbytes + b_bytes, bcol + TC_BK, n0);
                            } else {
                                tma_load_2d_u(tmb, full_b_u + 8u * sb, b_ring_u + (uint32_t)sb * b_sbytes, bcol, n0);
                                if (nsub == 2)
                                    tma_load_2d_u(tmb, full_b_u + 8u * sb, b_ring_u + (uint32_t)sb * b_sbytes + b_bytes, bcol + TC_BK, n0);
                            }
                        }
                        if (++sb == NSB) { sb = 0; pb ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
      if (rank == 0) {                                           // PAIR: the leader issues for both tensor cores
        auto umma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t i, uint32_t acc_) {
            if constexpr (PAIR) umma_bf16_pair(d, a, b, i, acc_); else umma_bf16(d, a, b, i, acc_);
        };
        auto commit = [](uint32_t bar) { if constexpr (PAIR) umma_commit_pair(bar); else umma_commit_u(bar); };
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;                                 // parity to wait for on the full barriers
        uint32_t acc = 0, pacc = 1;                              // accumulator buffer and parity of its empty barrier
        const bool dbg_on = T3_DBG_ON(blockIdx.x == (gridDim.x / 2 & ~1u) && elect_one()); (void)dbg_on;
        int dti = 0;
        const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
        const uint32_t full_a_u = smem_u32(&full_a[0]), empty_a_u = smem_u32(&empty_a[0]);
        const uint32_t full_b_u = smem_u32(&full_b[0]), empty_b_u = smem_u32(&empty_b[0]);
        const uint32_t tfull_u = smem_u32(&tmem_full_bar[0]), tempty_u = smem_u32(&tmem_empty_bar[0]);
        const bool leader = elect_one();
        const uint32_t dmode = (uint32_t)p.desc_mode;
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const T3Tile tt = t3_decode(p, tile, n_res);
            const uint32_t idesc = make_idesc(TILE_M, tt.bn);
            T3_CLK(8 * dti);
            mbar_wait_u(tempty_u + 8u * acc, pacc);              // epilogue has drained this accumulator
            tc_fence_after();
            T3_CLK(8 * dti + 1);
#ifdef CSI_T3_DEBUG
            bool dfirst = true;
#endif
            const uint32_t tacc = tmem_base + acc * 256u;
            uint32_t accum = 0;
            for (int gi = 0; gi < plan.ng; ++gi) {
                const int g_klen = plan.g[gi].klen, tap0 = plan.g[gi].tap0, ntaps = plan.g[gi].ntaps;
                for (int k0 = 0; k0 < g_klen; k0 += KSUB * TC_BK) {
                    const int nsub = (KSUB == 2 && g_klen - k0 > TC_BK) ? 2 : 1;
                    mbar_wait_u(full_a_u + 8u * sa, pa);
#ifdef CSI_T3_DEBUG
                    if (dfirst) { T3_CLK(8 * dti + 2); dfirst = false; }
#endif
                    const uint32_t a_addr = a_ring_u + (uint32_t)sa * a_sbytes;
                    const int rem_last = g_klen - k0 - (nsub - 1) * TC_BK;       // channels of the last block of this stage
                    const bool last_full = rem_last >= TC_BK;
                    const int last_steps = last_full ? 4 : (rem_last >> 4);
                    for (int t = 0; t < ntaps; ++t) {
                        if (tt.n0 >= plan.t[tap0 + t].n_hi || tt.n0 + tt.bn <= plan.t[tap0 + t].n_lo) continue;   // (same rule as the producer)
                        if (!resident || tile == tile_first) mbar_wait_u(full_b_u + 8u * sb, pb);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t roff = (uint32_t)plan.t[tap0 + t].a_row_off;
                            uint64_t adesc = make_kmajor_desc_bo(a_addr + roff * 128u, dmode ? roff : 0u);
                            uint64_t bdesc = make_kmajor_desc(b_ring_u + (uint32_t)sb * b_sbytes);
                            if (nsub == 2) {                        // first block of a two-block stage is always full
                                umma(tacc, adesc, bdesc, idesc, accum);
                                umma(tacc, adesc + 2, bdesc + 2, idesc, 1u);
                                umma(tacc, adesc + 4, bdesc + 4, idesc, 1u);
                                umma(tacc, adesc + 6, bdesc + 6, idesc, 1u);
                                adesc = make_kmajor_desc_bo(a_addr + a_bytes + roff * 128u, dmode ? roff : 0u);
                                bdesc = make_kmajor_desc(b_ring_u + (uint32_t)sb * b_sbytes + b_bytes);
                                accum = 1;
                            }
                            if (last_full) {
                                umma(tacc, adesc, bdesc, idesc, accum);
                                umma(tacc, adesc + 2, bdesc + 2, idesc, 1u);
                                umma(tacc, adesc + 4, bdesc + 4, idesc, 1u);
                                umma(tacc, adesc + 6, bdesc + 6, idesc, 1u);
                            } else {
                                for (int k = 0; k < last_steps; ++k)
                                    umma(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (accum | (uint32_t)k) ? 1u : 0u);
                            }
                            if (!resident) commit(empty_b_u + 8u * sb);
                        }
                        accum = 1;
                        if (++sb == NSB) { sb = 0; pb ^= 1u; }
                    }
                    if (leader) commit(empty_a_u + 8u * sa);
                    if (++sa == NSA) { sa = 0; pa ^= 1u; }
                }
            }
            if (leader) commit(tfull_u + 8u * acc);
            T3_CLK(8 * dti + 3);
            ++dti;
            acc ^= 1u;
            if (acc == 0) pacc ^= 1u;
        }
      }
    } else {
        // 8 epilogue warps: warp w drains TMEM lane quarter w % 4; the two warps of a quarter take alternate panels
        const int q = warp & 3, ew = warp - 2, half = ew >> 2;
        const int et = threadIdx.x - 64;                            // 0..255 among the epilogue threads
        const bool drop = p.drop_p > 0.f;
        DropCtx dc;
        if (drop) dc = drop_ctx(p.rng, p.drop_p);
        const int ld8 = ((p.N + 15) & ~15) >> 3;
        uint8_t* mybuf = cstage + (size_t)ew * p.epi_bufs * 4096;
        const bool two_bufs = p.epi_bufs == 2;
        int ti = 0, sbuf = 0;
        const bool dbg_on = T3_DBG_ON(blockIdx.x == (gridDim.x / 2 & ~1u) && ew == 0 && lane == 0); (void)dbg_on;
        const uint32_t tempty_remote = PAIR ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0u) : 0u;
        // RES: the fp32 residual panels (32 rows x 128 B, the geometry of the output panels) of this warp arrive by TMA in a
        // two-slot ring that runs TWO panels ahead of the arithmetic, across tile boundaries: 64 KB in flight per SM.  With
        // register prefetch one panel ahead (32 KB per SM) the epilogue sat on the load latency: out-proj 20 -> 37 us.
        uint8_t* rbuf = cstage + (size_t)(T3_EPI_THREADS / 32) * p.epi_bufs * 4096 + (size_t)ew * 2 * 4096;
        int pf_tile = tile_first, pf_pi = half;
        uint32_t pf_cnt = 0, rcc = 0;
        auto pf_issue = [&]() {
            if constexpr (!RES) return;
            while (pf_tile < tile_end) {
                const T3Tile t = t3_decode(p, pf_tile, n_res);
                if (pf_pi < (t.bn + PW - 1) / PW) {
                    if (lane == 0) {
                        const int pc0 = pf_pi * PW, w = min(PW, t.bn - pc0);
                        const uint32_t slot = pf_cnt & 1u, bar = smem_u32(&res_full[ew][slot]);
                        mbar_expect_tx_u(bar, (uint32_t)(32 * w) * 4u);
                        tma_load_2d_u(w == PW ? &tmR : &tmRt, bar, smem_u32(rbuf + slot * 4096), t.n0 + pc0,
                                      t.mt * TILE_M + (int)rank * TC_BM + q * 32);
                    }
                    ++pf_cnt; pf_pi += 2;
                    return;
                }
                pf_tile += tile_step; pf_pi = half;
            }
        };
        if constexpr (RES) { pf_issue(); pf_issue(); }
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++ti) {
            const int acc = ti & 1;
            const uint32_t aph = (uint32_t)(ti >> 1) & 1u;
            const T3Tile tt = t3_decode(p, tile, n_res);
            const int m0 = tt.mt * TILE_M + (int)rank * TC_BM, n0 = tt.n0, bn = tt.bn;
            const int npan = (bn + PW - 1) / PW;
            const int m = m0 + q * 32 + lane;
            if (p.bias) {
                sbias[acc][et] = (et < bn && (n0 + et) < p.N) ? p.bias[n0 + et] : 0.f;
                named_bar_sync(1, T3_EPI_THREADS);
            }
            T3_CLK(1000 + 8 * ti);
            mbar_wait(&tmem_full_bar[acc], aph);
            tc_fence_after();
            T3_CLK(1000 + 8 * ti + 1);
            const uint32_t tacc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
            for (int pi = half; pi < npan; pi += 2) {
                const int pc0 = pi * PW;
                const int width = min(PW, bn - pc0);                // multiple of 16
                float v[PW];
                uint32_t r[PW];
                // all TMEM loads of the panel are issued before the single wait
#pragma unroll
                for (int c = 0; c < PW; c += 16) {
                    if (c < width) tmem_ld16(tacc + (uint32_t)(pc0 + c), r + c);
                    else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[c + j] = 0u;
                    }
                }
                tmem_ld_wait();
                if (pi == half) T3_CLK(1000 + 8 * ti + 2);
#pragma unroll
                for (int j = 0; j < PW; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < PW; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&sbias[acc][pc0 + j]);
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
                if (drop) {
#pragma unroll
                    for (int g8 = 0; g8 < PW / 8; ++g8) {
                        float ks[8];
                        drop_scales8(dc, p.drop_site, (unsigned long long)m * ld8 + ((n0 + pc0) >> 3) + g8, ks);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[g8 * 8 + j] *= ks[j];
                    }
                }
                if constexpr (RES) {
                    const uint32_t slot = rcc & 1u;
                    mbar_wait(&res_full[ew][slot], (rcc >> 1) & 1u);
                    const uint8_t* rb = rbuf + slot * 4096;
                    if (width == PW) {
#pragma unroll
                        for (int c16 = 0; c16 < 8; ++c16) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rb + lane * 128 + ((c16 ^ (lane & 7)) << 4));
                            v[c16 * 4] += r4.x; v[c16 * 4 + 1] += r4.y; v[c16 * 4 + 2] += r4.z; v[c16 * 4 + 3] += r4.w;
                        }
                    } else {
                        const uint8_t* rowp = rb + lane * (width * 4);
#pragma unroll
                        for (int c16 = 0; c16 < 8; ++c16) {
                            if (c16 * 4 < width) {
                                const float4 r4 = *reinterpret_cast<const float4*>(rowp + (c16 << 4));
                                v[c16 * 4] += r4.x; v[c16 * 4 + 1] += r4.y; v[c16 * 4 + 2] += r4.z; v[c16 * 4 + 3] += r4.w;
                            }
                        }
                    }
                    __syncwarp();                                    // every lane has read its row: refill the slot
                    pf_issue();
                    ++rcc;
                }
                // stage the 32-row slice of this warp (row = lane) and bulk-store it; the tensor maps clip rows >= M and
                // columns >= N.  Full panels are 128-byte rows in the TMA 128B swizzle; the narrower last panel of a tile
                // uses linear rows of width*es bytes and its own (unswizzled) tensor map.
                if (lane == 0) { if (two_bufs) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }   // the buffer about to be written has been read
                __syncwarp();
                if (pi == half) T3_CLK(1000 + 8 * ti + 3);
                uint8_t* buf = mybuf + (size_t)sbuf * 4096;
                if (width == PW) {
                    uint8_t* rowp = buf + lane * 128;
#pragma unroll
                    for (int c16 = 0; c16 < 8; ++c16) {
                        uint4 u;
                        if constexpr (sizeof(TC) == 2) {
                            u.x = pack2_bf16(v[c16 * 8 + 0], v[c16 * 8 + 1]); u.y = pack2_bf16(v[c16 * 8 + 2], v[c16 * 8 + 3]);
                            u.z = pack2_bf16(v[c16 * 8 + 4], v[c16 * 8 + 5]); u.w = pack2_bf16(v[c16 * 8 + 6], v[c16 * 8 + 7]);
                        } else {
                            u.x = __float_as_uint(v[c16 * 4 + 0]); u.y = __float_as_uint(v[c16 * 4 + 1]);
                            u.z = __float_as_uint(v[c16 * 4 + 2]); u.w = __float_as_uint(v[c16 * 4 + 3]);
                        }
                        *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) = u;
                    }
                } else {
                    const int pitch = width * (int)sizeof(TC);
                    uint8_t* rowp = buf + lane * pitch;
#pragma unroll
                    for (int c16 = 0; c16 < 8; ++c16) {
                        if (c16 * 16 < pitch) {
                            uint4 u;
                            if constexpr (sizeof(TC) == 2) {
                                u.x = pack2_bf16(v[c16 * 8 + 0], v[c16 * 8 + 1]); u.y = pack2_bf16(v[c16 * 8 + 2], v[c16 * 8 + 3]);
                                u.z = pack2_bf16(v[c16 * 8 + 4], v[c16 * 8 + 5]); u.w = pack2_bf16(v[c16 * 8 + 6], v[c16 * 8 + 7]);
                            } else {
                                u.x = __float_as_uint(v[c16 * 4 + 0]); u.y = __float_as_uint(v[c16 * 4 + 1]);
                                u.z = __float_as_uint(v[c16 * 4 + 2]); u.w = __float_as_uint(v[c16 * 4 + 3]);
                            }
                            *reinterpret_cast<uint4*>(rowp + (c16 << 4)) = u;
                        }
                    }
                }
                if (pi == half) T3_CLK(1000 + 8 * ti + 6);
                fence_proxy_async();
                __syncwarp();
                if (pi == half) T3_CLK(1000 + 8 * ti + 7);
                if (lane == 0) {
                    if (width == PW) tma_store_2d(&tmC, buf, n0 + pc0, m0 + q * 32);
                    else tma_store_2d(&tmCt, buf, n0 + pc0, m0 + q * 32);
                    tma_store_commit();
                }
                if (pi == half) T3_CLK(1000 + 8 * ti + 4);
                if (two_bufs) sbuf ^= 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(tempty_remote + 8u * (uint32_t)acc);
                else mbar_arrive(&tmem_empty_bar[acc]);
            }
            T3_CLK(1000 + 8 * ti + 5);
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

static int g_force_ntn = 0;          // > 0: split N into this many column tiles (A/B runs of the tile-shape heuristic)
extern "C" int csi_set_gemm_ntn(int ntn) { g_force_ntn = ntn > 0 ? ntn : 0; return CSI_OK; }

static int bn_for(int N, int ntn) {
    int bn = ((N + ntn - 1) / ntn + 15) & ~15;
    return bn < 16 ? 16 : bn;
}
static int pick_bn3(int N, int mtiles, int sms) {
    const int base = (N + 255) / 256;                  // fewest column tiles that fit one accumulator (<= 256 columns)
    if (g_force_ntn > 0) return bn_for(N, g_force_ntn < base ? base : g_force_ntn);
    (void)mtiles; (void)sms;
    return bn_for(N, base);
}

static int g_num_sms3 = 0;
static int g_desc_mode = 0;       // 0: base-offset field left 0 (swizzle follows the absolute address); 1: base offset = row % 8
static int g_tap_share = 1;
static int g_pdl3 = -1;            // programmatic dependent launch: off by default (measured in-step: 4.79 ms with, 4.78 ms without), CSI_PDL=1 enables
static long long* g_t3_dbg = nullptr;
extern "C" int csi_set_gemm_debug(long long* buf) { g_t3_dbg = buf; return CSI_OK; }
static int g_resident = -1;        // weight-stationary short-K mode; CSI_GEMM_RESIDENT=0 or csi_set_gemm_resident(0) disables
extern "C" int csi_set_gemm_resident(int on) { g_resident = on ? 1 : 0; return CSI_OK; }
static int g_pair = -1;            // CTA pairs (cta_group::2) for the NT GEMM; CSI_GEMM_2CTA=0 or csi_set_gemm_pair(0) = single-CTA tiles
extern "C" int csi_set_gemm_pair(int mode) { g_pair = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return CSI_OK; }
extern "C" int csi_set_gemm_pdl(int on) { g_pdl3 = on ? 1 : 0; return CSI_OK; }
extern "C" int csi_set_gemm_desc_mode(int mode) { g_desc_mode = mode ? 1 : 0; return CSI_OK; }
extern "C" int csi_set_gemm_tap_share(int on) { g_tap_share = on ? 1 : 0; return CSI_OK; }

extern "C" int csi_gemm_nt_tc3_banded(const void* A, int lda, const void* Bw, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                               const csi_seg* segs, const csi_band* bands, int nseg, const float* bias, const float* residual, int ldr,
                               float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream);

// row-box map for A: box = box_rows x 64 columns (box_rows <= 256)
extern "C" int csi_gemm_nt_tc3(const void* A, int lda, const void* Bw, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                               const csi_seg* segs, int nseg, const float* bias, const float* residual, int ldr,
                               float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    return csi_gemm_nt_tc3_banded(A, lda, Bw, ldb, C, ldc, c_dtype, M, N, segs, nullptr, nseg, bias, residual, ldr, drop_p, drop_site,
                                  rng, stream);
}

// bands[s] = {n_lo, n_hi}: segment s contributes to output columns [n_lo, n_hi) only (its weights are zero elsewhere).
// NULL = every segment contributes everywhere.
extern "C" int csi_gemm_nt_tc3_banded(const void* A, int lda, const void* Bw, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                               const csi_seg* segs, const csi_band* bands, int nseg, const float* bias, const float* residual, int ldr,
                               float drop_p, unsigned drop_site, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(A && Bw && C && segs, "null pointer");
    CSI_CHECK_ARG(!(drop_p > 0.f) || rng, "dropout needs rng");
    CSI_CHECK_ARG(nseg >= 1 && nseg <= CSI_MAX_SEGS, "1..32 segments");
    const int es = c_dtype == CSI_BF16 ? 2 : 4;
    CSI_CHECK_ARG((ldc * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "C must be 16-byte aligned with a 16-byte row pitch");
    CSI_CHECK_ARG(!residual || (ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0), "residual must be 16-byte aligned");
    CSI_CHECK_ARG(!(residual && c_dtype == CSI_BF16), "residual is only fused for fp32 output");
    // ---- group consecutive segments that read the same A columns at row shifts within a 16-row window
    T3Plan plan;
    plan.ng = 0;
    int min_shift = 0, max_shift = 0, a_cols = 0, b_cols = 0, span = 0;
    for (int i = 0; i < nseg; ++i) {
        const csi_seg s = segs[i];
        if (s.a_row_shift < min_shift) min_shift = s.a_row_shift;
        if (s.a_row_shift > max_shift) max_shift = s.a_row_shift;
        if (s.a_col_off + s.klen > a_cols) a_cols = s.a_col_off + s.klen;
        if (s.b_col_off + s.klen > b_cols) b_cols = s.b_col_off + s.klen;
        bool joined = false;
        if (g_tap_share && plan.ng > 0) {
            T3Group& g = plan.g[plan.ng - 1];
            if (g.a_col_off == s.a_col_off && g.klen == s.klen) {
                int lo = g.min_shift, hi = g.min_shift;
                for (int t = 0; t < g.ntaps; ++t) {
                    const int sh = g.min_shift + plan.t[g.tap0 + t].a_row_off;
                    if (sh > hi) hi = sh;
                }
                const int nlo = s.a_row_shift < lo ? s.a_row_shift : lo, nhi = s.a_row_shift > hi ? s.a_row_shift : hi;
                if (nhi - nlo <= 16) {
                    if (nlo != lo)
                        for (int t = 0; t < g.ntaps; ++t) plan.t[g.tap0 + t].a_row_off = (short)(plan.t[g.tap0 + t].a_row_off + (lo - nlo));
                    g.min_shift = nlo;
                    plan.t[g.tap0 + g.ntaps].a_row_off = (short)(s.a_row_shift - nlo);
                    plan.t[g.tap0 + g.ntaps].pad = 0;
                    plan.t[g.tap0 + g.ntaps].b_col_off = s.b_col_off;
                    plan.t[g.tap0 + g.ntaps].n_lo = bands ? bands[i].n_lo : 0;
                    plan.t[g.tap0 + g.ntaps].n_hi = bands ? bands[i].n_hi : 0x7fffffff;
                    ++g.ntaps;
                    if (nhi - nlo > span) span = nhi - nlo;
                    joined = true;
                }
            }
        }
        if (!joined) {
            T3Group& g = plan.g[plan.ng];
            g.a_col_off = s.a_col_off; g.klen = s.klen; g.min_shift = s.a_row_shift; g.tap0 = i; g.ntaps = 1;
            plan.t[i].a_row_off = 0; plan.t[i].pad = 0; plan.t[i].b_col_off = s.b_col_off;
            plan.t[i].n_lo = bands ? bands[i].n_lo : 0;
            plan.t[i].n_hi = bands ? bands[i].n_hi : 0x7fffffff;
            ++plan.ng;
        }
    }
    CSI_CHECK_ARG(a_cols <= lda && b_cols <= ldb, "segment exceeds the row pitch");
    if (g_num_sms3 == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_num_sms3, cudaDevAttrMultiProcessorCount, dev));
    }
    if (g_pair < 0) { const char* e = getenv("CSI_GEMM_2CTA"); g_pair = (e && e[0] == '0') ? 0 : 1; }
    // Pairs pay off where the mainloop dominates (shared-memory bound: conv / data-gradient shapes, 753 -> 910 TF/s at k=5);
    // short-K calls are bound by their epilogue stores and lose a little to the lock-step of the two CTAs (QKV: 35 -> 39 us),
    // so they keep single-CTA tiles.  csi_set_gemm_pair(2) forces pairs for every shape (tests).
    int ksum = 0;
    for (int i = 0; i < nseg; ++i) ksum += segs[i].klen;
    bool pair = g_num_sms3 >= 2 && (g_pair == 2 || (g_pair == 1 && ksum >= 512));
    // ---- banded call: column tiles follow the band boundaries (each column group between two boundaries is cut like a GEMM
    //      of its own), so that no tile mixes columns with different tap sets
    int ntab = 0, tab_n0[T3_MAX_NTAB], tab_bn[T3_MAX_NTAB], bn_max = 0, tab_tail = 0;
    if (bands) {
        int cuts[2 * CSI_MAX_SEGS + 2], nc = 0;
        cuts[nc++] = 0; cuts[nc++] = N;
        bool ok = true;
        for (int i = 0; i < nseg; ++i) {
            CSI_CHECK_ARG(bands[i].n_lo >= 0 && bands[i].n_lo < bands[i].n_hi, "empty band");
            const int lo = bands[i].n_lo, hi = bands[i].n_hi < N ? bands[i].n_hi : N;
            if (lo % 16 || (hi % 16 && hi != N)) ok = false;
            cuts[nc++] = lo; cuts[nc++] = hi;
        }
        for (int i = 1; i < nc; ++i)                          // insertion sort + unique
            for (int j = i; j > 0 && cuts[j] < cuts[j - 1]; --j) { const int tmp = cuts[j]; cuts[j] = cuts[j - 1]; cuts[j - 1] = tmp; }
        int nu = 1;
        for (int i = 1; i < nc; ++i) if (cuts[i] != cuts[nu - 1]) cuts[nu++] = cuts[i];
        for (int gI = 0; ok && gI + 1 < nu; ++gI) {
            const int w = cuts[gI + 1] - cuts[gI], nt = (w + 255) / 256, bn = bn_for(w, nt);
            for (int t = 0, n0 = cuts[gI]; t < nt; ++t, n0 += bn) {
                if (ntab == T3_MAX_NTAB) { ok = false; break; }
                int wt = cuts[gI + 1] - n0;
                if (wt > bn) wt = bn;
                tab_n0[ntab] = n0; tab_bn[ntab] = (wt + 15) & ~15;
                if (tab_bn[ntab] > bn_max) bn_max = tab_bn[ntab];
                ++ntab;
            }
        }
        // the narrower last store panel of a tile goes through ONE extra tensor map: all tiles must agree on its width
        const int pw = 128 / (c_dtype == CSI_BF16 ? 2 : 4);
        int tailw = 0;
        for (int i = 0; ok && i < ntab; ++i) {
            const int tw = tab_bn[i] % pw;
            if (tw && tailw && tw != tailw) ok = false;
            if (tw) tailw = tw;
        }
        if (!ok) ntab = 0;                                     // boundaries off the 16-column grid: uniform tiles, taps still skipped per tile
        else tab_tail = tailw;
    }
    const int BN = ntab ? bn_max : pick_bn3(N, (M + TC_BM - 1) / TC_BM, g_num_sms3);
    const int a_rows = TC_BM + ((span + 7) & ~7);
    static int ksub_env = -1;                            // CSI_GEMM_KSUB=1 forces one 64-channel block per stage (A/B runs)
    if (ksub_env < 0) { const char* e = getenv("CSI_GEMM_KSUB"); ksub_env = (e && e[0] == '1') ? 1 : 2; }
    // One staging buffer per epilogue warp (two were measured: no difference) leaves 32 KB more for the operand rings.
    static int epi_bufs = 0;                             // CSI_GEMM_EPIBUF=2 restores double-buffered staging (A/B runs)
    if (!epi_bufs) { const char* e = getenv("CSI_GEMM_EPIBUF"); epi_bufs = (e && e[0] == '2') ? 2 : 1; }
    const size_t fixed = 1024 + 8 * (size_t)epi_bufs * 4096 + (residual ? 8 * 2 * 4096 : 0);   // + the residual panel rings
    const size_t budget = 224 * 1024 - fixed;            // + 3 KB of static shared memory = the 227 KB an SM offers
    // ---- weight-stationary mode (Nt3Params::resident): a CTA pair keeps all weight stages of its column tile in shared
    //      memory when they leave room for at least four A stages
    if (g_resident < 0) { const char* e = getenv("CSI_GEMM_RESIDENT"); g_resident = (e && e[0] == '0') ? 0 : 1; }
    int resident = 0, res_ksub = 1, res_nst = 0, res_nsa = 0;
    if (g_resident && g_pair != 0 && g_num_sms3 >= 2 && !bands) {
        res_ksub = (ksub_env == 2 && BN <= 160) ? 2 : 1;
        for (int gi = 0; gi < plan.ng; ++gi)
            res_nst += ((plan.g[gi].klen + res_ksub * TC_BK - 1) / (res_ksub * TC_BK)) * plan.g[gi].ntaps;
        const size_t res_b = (size_t)res_nst * (BN / 2) * 128 * res_ksub, a_st = (size_t)a_rows * 128 * res_ksub;
        if (res_nst <= T3_MAX_STAGES && res_b + 4 * a_st <= budget) {
            resident = 1;
            pair = true;
            res_nsa = (int)((budget - res_b) / a_st);
            if (res_nsa > T3_MAX_STAGES) res_nsa = T3_MAX_STAGES;
        }
    }
    const int tile_m = pair ? 2 * TC_BM : TC_BM, units = pair ? g_num_sms3 / 2 : g_num_sms3;   // schedulable units: CTAs or CTA pairs
    CUtensorMap tmA, tmB, tmC;
    const bf16* a_base = reinterpret_cast<const bf16*>(A) + (long long)min_shift * lda;
    int rc = make_map(&tmA, a_base, (long long)M + (max_shift - min_shift), a_cols, lda, a_rows);
    if (rc) return rc;
    rc = make_map(&tmB, Bw, N, b_cols, ldb, pair ? BN / 2 : BN);      // a CTA of a pair stages half of the weight tile
    if (rc) return rc;
    rc = make_map_ex(&tmC, C, M, N, ldc, 32, 128 / es, es);
    if (rc) return rc;
    const int tail = ntab ? tab_tail : BN % (128 / es);   // last panel of a tile when its width is not a whole number of panels
    CUtensorMap tmCt = tmC;
    if (tail) {
        rc = make_map_ex(&tmCt, C, M, N, ldc, 32, tail, es, false);
        if (rc) return rc;
    }
    CUtensorMap tmR = tmC, tmRt = tmC;                   // fp32 residual: the geometry of the output panels
    if (residual) {
        rc = make_map_ex(&tmR, residual, M, N, ldr, 32, 32, 4);
        if (rc) return rc;
        tmRt = tmR;
        if (tail) {
            rc = make_map_ex(&tmRt, residual, M, N, ldr, 32, tail, 4, false);
            if (rc) return rc;
        }
    }
    Nt3Params p;
    p.M = M; p.N = N; p.BN = BN; p.C = C; p.ldc = ldc; p.a_rows = a_rows;
    p.ntn = ntab ? ntab : (N + BN - 1) / BN;
    p.ntab = ntab;
    for (int i = 0; i < ntab; ++i) { p.tab_n0[i] = (short)tab_n0[i]; p.tab_bn[i] = (short)tab_bn[i]; }
    const int mtiles = (M + tile_m - 1) / tile_m;
    p.ntiles = p.ntn * mtiles;
    // tail-wave split (see Nt3Params): 308 row tiles on 148 SMs are 2 full waves + 12 tiles; as 12 x 5 pieces of 64 columns
    // the third wave costs a fraction of a tile instead of a whole one
    static int tail_split = -1;                          // CSI_GEMM_TAILSPLIT=0 disables (A/B runs)
    if (tail_split < 0) { const char* e = getenv("CSI_GEMM_TAILSPLIT"); tail_split = (e && e[0] == '0') ? 0 : 1; }
    p.nfull = p.ntiles; p.npiece = 1; p.BN2 = BN; p.mt_tail = mtiles;
    CUtensorMap tmB2 = tmB;
    p.resident = resident; p.mtiles = mtiles; p.dbg = g_t3_dbg;
    if (tail_split && !resident && !ntab && p.ntn == 1 && mtiles > units && BN > 64) {
        const int leftover = mtiles % units, npiece = (BN + 63) / 64;
        if (leftover > 0 && leftover * npiece <= units) {
            rc = make_map(&tmB2, Bw, N, b_cols, ldb, pair ? 32 : 64);
            if (rc) return rc;
            p.nfull = mtiles - leftover; p.mt_tail = p.nfull; p.npiece = npiece; p.BN2 = 64;
            p.ntiles = p.nfull + leftover * npiece;
        }
    }
    p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.drop_p = drop_p; p.drop_site = drop_site; p.rng = rng;
    p.row_base = -min_shift;
    p.desc_mode = g_desc_mode;
    if (g_pdl3 < 0) { const char* e = getenv("CSI_PDL"); g_pdl3 = (e && e[0] == '1') ? 1 : 0; }
    p.pdl = g_pdl3;
    p.epi_bufs = epi_bufs;
    int ksub = (ksub_env == 2 && BN <= 160) ? 2 : 1, nsa = 0, nsb = 0;
    size_t a_bytes = 0, b_bytes = 0;                     // per ring stage
    for (;; ksub = 1) {
        a_bytes = (size_t)a_rows * 128 * ksub;
        b_bytes = (size_t)(pair ? BN / 2 : BN) * 128 * ksub;
        if (span == 0) {                               // every group is a single tap: A and B stages pair up
            nsa = (int)(budget / (a_bytes + b_bytes));
            if (nsa > T3_MAX_STAGES) nsa = T3_MAX_STAGES;
            nsb = nsa;
        } else {
            nsa = ksub == 2 ? 2 : 3;
            nsb = (int)((budget - nsa * a_bytes) / b_bytes);
            if (nsb > T3_MAX_STAGES) nsb = T3_MAX_STAGES;
        }
        if (ksub == 1 || nsb >= 3) break;                // two-block stages need a ring of at least three
    }
    if (nsa < 2) nsa = 2;
    if (nsb < 2) nsb = 2;
    if (resident) {
        ksub = res_ksub; nsa = res_nsa; nsb = res_nst;
        a_bytes = (size_t)a_rows * 128 * ksub;
        b_bytes = (size_t)(BN / 2) * 128 * ksub;
    }
    p.ksub = ksub;
    p.nsa = nsa; p.nsb = nsb;
    const size_t smem = fixed + nsa * a_bytes + nsb * b_bytes;
    int grid = units;
    if (p.ntiles < grid) grid = p.ntiles;
    if (resident && grid < p.ntn) grid = p.ntn;             // every column tile needs an owner
    if (pair) grid *= 2;
#define LAUNCH3(TC, RES, PAIR)                                                                                          \
    do {                                                                                                                \
        CSI_CUDA(cudaFuncSetAttribute(gemm_nt_tc3_kernel<TC, RES, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cudaLaunchConfig_t cfg = {};                                                                                    \
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T3_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ST(stream); \
        cudaLaunchAttribute at[2];                                                                                      \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                  \
        at[0].val.programmaticStreamSerializationAllowed = p.pdl;                                                       \
        at[1].id = cudaLaunchAttributeClusterDimension;                                                                 \
        at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;                             \
        cfg.attrs = at; cfg.numAttrs = PAIR ? 2 : 1;                                                                    \
        CSI_CUDA(cudaLaunchKernelEx(&cfg, gemm_nt_tc3_kernel<TC, RES, PAIR>, tmA, tmB, tmB2, tmC, tmCt, tmR, tmRt, p, plan)); \
    } while (0)
#define LAUNCH3P(TC, RES) do { if (pair) LAUNCH3(TC, RES, true); else LAUNCH3(TC, RES, false); } while (0)
    if (c_dtype == CSI_BF16) LAUNCH3P(bf16, false);
    else if (residual) LAUNCH3P(float, true);
    else LAUNCH3P(float, false);
#undef LAUNCH3P
#undef LAUNCH3
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
