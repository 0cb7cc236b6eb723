// tcgen05 / TMEM attention backward (that.py:149 under autograd), bf16 activations.  See attention_tc.cu for the layout
// conventions (head groups sharing 128-byte lines, descriptors advanced inside the swizzle atom).
//
// Persistent CTAs walk the items (sample, head group).  Warp roles: 16 worker warps (4 per TMEM lane quarter: the query
// columns of a block are split between them), one MMA-issue warp, one TMA loader warp (Q, K, V, dO boxes of the next item
// land while the current one is processed).
//
// TRANSPOSED formulation -- a thread owns a KEY row -- so that the probabilities can stay in TMEM as MMA operands:
//   per head, per 128-key tile j, per query block qb (<= 256 queries, 64-aligned):
//     S^T = K_j Q_qb^T,  dP^T = V_j dO_qb^T                      tcgen05.mma, M=128 keys, N = block, K = hp     -> TMEM
//     P^T = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - D[q])       workers; bf16 pairs written back IN PLACE (each warp packs
//                                                                 into the first half of its own column range), and dS^T
//                                                                 once more to shared memory ([key][query], 128B swizzle)
//     dV_j (+)= P^T dO_qb,  dK_j (+)= dS^T Q_qb                   A from TMEM, B = dO / Q read in place as MN-major operands
//     dQ_qb (+)= dS K_j                                           A = the shared dS^T tile read MN-major, B = K_j MN-major
//   dV_j / dK_j are drained after the last query block, dQ after the last key tile; the in_proj bias gradient (column
//   sums of the stored bf16 dq | dk | dv) is reduced per warp with shuffles, per CTA in shared memory, and added to
//   global memory with one atomic per channel and item.
#include "tc_common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define AB_LOG2E 1.4426950408889634f
#define AB_WORKER_WARPS 16
#define AB_THREADS (32 * (AB_WORKER_WARPS + 2))
#define AB_MAX_QB 4
#define AB_MAX_MT 8

__device__ __forceinline__ uint64_t ab_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float ab_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t ab_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void ab_tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ab_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ab_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ab_sts128(uint32_t saddr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void ab_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct AbParams {
    int B, L, Lk, Lp, H, hp, hd, G, NG, halo;
    int RB, nbox;
    int nkt;                                        // 128-key tiles per head
    int nqb, qb_n0[AB_MAX_QB], qb_len[AB_MAX_QB], qb_mt0[AB_MAX_QB], qb_nmt[AB_MAX_QB];   // query blocks and their dQ M-tiles
    // dQ M-tile t reads the 128 queries starting at 64-query block mt_skb[t] of the shared dS^T tile (the last tile of a
    // head is shifted back so that it stays inside the tile) and owns the queries [mt_q0[t], mt_q1[t])
    int mt_skb[AB_MAX_MT], mt_q0[AB_MAX_MT], mt_q1[AB_MAX_MT];
    int ldo, lddq, HP, d;
    const bf16* o;
    bf16* dqkv;
    const float* lse;
    float* dbias;
    float sc;
    uint32_t opnd_bytes, stage_bytes, dst_bytes;    // one operand box set, one stage (Q|K|V|dO), the shared dS^T tile
    int nstage;
    int dp_col, dv_col, dk_col, dq_col;             // TMEM columns (S^T sits at 0)
    long long* dbg;                                 // optional phase clocks of one CTA, NULL in production
};

// column sums over the 32 rows (lanes) of 16 per-lane values: butterfly transpose-reduce, 15 + 1 shuffles.
// Returns in every lane the total of column ab_col16(lane).
__device__ __forceinline__ float ab_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const float give = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const float give = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const float give = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    {
        const bool up = lane & 2;
        const float give = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int ab_col16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// 16 accumulator columns of this lane's row -> scale -> bf16 -> global row (if valid); the rounded values' column sums -> csum
__device__ __forceinline__ void ab_drain16(uint32_t taddr, float scale, bool valid, bf16* dst, float* csum, int lane) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_ld_wait();
    float v[16];
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        pk[j] = ab_pack(__uint_as_float(r[2 * j]) * scale, __uint_as_float(r[2 * j + 1]) * scale);
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
        v[2 * j] = valid ? __low2float(h) : 0.f;
        v[2 * j + 1] = valid ? __high2float(h) : 0.f;
    }
    if (valid) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (csum) {
        const float t = ab_colsum16(v, lane);
        if (!(lane & 1)) atomicAdd(csum + ab_col16(lane), t);
    }
}

#define AB_CLK(slot) do { if (dbg_on && (slot) < 1000) p.dbg[(slot)] = clock64(); } while (0)
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                    const __grid_constant__ CUtensorMap tmDO, AbParams p) {
    extern __shared__ __align__(1024) uint8_t ab_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~(uintptr_t)1023);
    // [stages: Q | K | V | dO][dS^T tile][lse2: G x Lk][D: G x Lk][csum: 192][barriers]
    uint8_t* dst_tile = smem + (size_t)p.nstage * p.stage_bytes;
    float* lse2 = reinterpret_cast<float*>(dst_tile + p.dst_bytes);
    float* Dv = lse2 + p.G * p.Lk;
    float* csum = Dv + p.G * p.Lk;
    uint64_t* bars = reinterpret_cast<uint64_t*>(csum + 192);
    uint64_t* full = bars;            // [2]  loader -> everyone
    uint64_t* empty = bars + 2;       // [2]  workers -> loader
    uint64_t* sdp_full = bars + 4;    // MMA -> workers: S^T and dP^T of a job are in TMEM
    uint64_t* pds_ready = bars + 5;   // workers -> MMA: P^T / dS^T written (16 warp arrivals)
    uint64_t* grad_done = bars + 6;   // MMA -> workers: gradient products of the last query block of a key tile completed
    uint64_t* drained = bars + 7;     // workers -> MMA: dV/dK (and dQ) accumulators have been read out (16 warp arrivals)
    uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = p.B * p.NG;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(sdp_full, 1);
        mbar_init(pds_ready, AB_WORKER_WARPS);
        mbar_init(grad_done, 1);
        mbar_init(drained, AB_WORKER_WARPS);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t dst_u = smem_u32(dst_tile);

    if (warp == AB_WORKER_WARPS + 1) {
        // ===== TMA loader (warp-uniform loop, lane 0 issues)
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it % p.nstage;
            const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
            mbar_wait(&empty[st], ph ^ 1u);
            if (lane == 0) {
                const int b = item / p.NG, g = item - b * p.NG;
                const int row0 = b * p.Lp + p.halo;
                mbar_expect_tx(&full[st], 4u * (uint32_t)p.Lk * 128u);
                uint8_t* dst = smem + (size_t)st * p.stage_bytes;
                for (int w = 0; w < 3; ++w)
                    for (int bx = 0; bx < p.nbox; ++bx)
                        tma_load_2d(&tmQKV, &full[st], dst + (size_t)w * p.opnd_bytes + (size_t)bx * p.RB * 128, w * p.HP + g * 64,
                                    row0 + bx * p.RB);
                for (int bx = 0; bx < p.nbox; ++bx)
                    tma_load_2d(&tmDO, &full[st], dst + (size_t)3 * p.opnd_bytes + (size_t)bx * p.RB * 128, g * 64, row0 + bx * p.RB);
            }
            __syncwarp();
        }
    } else if (warp == AB_WORKER_WARPS) {
        // ===== MMA issuer (warp-uniform loop, lane 0 issues)
        const uint32_t idesc_kk_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 4) << 24);        // K-major A and B, N added per block
        const uint32_t idesc_ts = make_idesc(TC_BM, p.hp) | (1u << 16);                                          // A from TMEM, B MN-major
        const uint32_t idesc_mm = make_idesc(TC_BM, p.hp) | (1u << 15) | (1u << 16);                             // A and B MN-major
        const int ksteps_h = p.hp >> 4;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0;
        int mj = 0;
        uint32_t ph_pds = 0, ph_dr = 0;
        int it = 0, ndrain_waits = 0;
        bool first_kt = true;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it % p.nstage;
            const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
            const int g = item % p.NG;
            const int nheads = min(p.G, p.H - g * p.G);
            mbar_wait(&full[st], ph);
            const uint32_t q_u = smem_u32(smem) + (uint32_t)st * p.stage_bytes, k_u = q_u + p.opnd_bytes, v_u = k_u + p.opnd_bytes,
                           do_u = v_u + p.opnd_bytes;
            for (int h = 0; h < nheads; ++h) {
                const uint32_t cb = (uint32_t)(h * p.hp * 2);
                for (int j = 0; j < p.nkt; ++j) {
                    const int kvalid = min(128, p.Lk - 128 * j);                  // key rows of this tile that were loaded
                    for (int qb = 0; qb < p.nqb; ++qb) {
                        const int n0 = p.qb_n0[qb], nlen = p.qb_len[qb];
                        AB_CLK(8 * mj + 0);
                        if (lane == 0) {
                            tc_fence_after();
                            // S^T = K_j Q_qb^T -> [0, nlen);  dP^T = V_j dO_qb^T -> [dp_col, dp_col + nlen)
                            const uint32_t idesc = idesc_kk_base | ((uint32_t)(nlen >> 3) << 17);
                            const uint64_t ka = make_kmajor_desc(k_u + (uint32_t)(128 * j) * 128u + cb), qd = make_kmajor_desc(q_u + (uint32_t)n0 * 128u + cb);
                            const uint64_t va = make_kmajor_desc(v_u + (uint32_t)(128 * j) * 128u + cb), gd = make_kmajor_desc(do_u + (uint32_t)n0 * 128u + cb);
#pragma unroll 4
                            for (int kk = 0; kk < ksteps_h; ++kk)
                                umma_bf16(tmem_base, ka + (uint64_t)(2 * kk), qd + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
#pragma unroll 4
                            for (int kk = 0; kk < ksteps_h; ++kk)
                                umma_bf16(tmem_base + (uint32_t)p.dp_col, va + (uint64_t)(2 * kk), gd + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
                            umma_commit(sdp_full);
                        }
                        __syncwarp();
                        // the accumulators of the previous key tile must have been read out before they are overwritten
                        if (qb == 0 && !first_kt) {
                            mbar_wait(drained, ph_dr);
                            ph_dr ^= 1u;
                            ++ndrain_waits;
                        }
                        first_kt = false;
                        AB_CLK(8 * mj + 1);
                        mbar_wait(pds_ready, ph_pds);
                        ph_pds ^= 1u;
                        AB_CLK(8 * mj + 2);
                        if (lane == 0) {
                            tc_fence_after();
                            // packed P^T / dS^T: worker part k packs its 16-column groups [gb, ge) into columns 16*gb + 8*(g - gb)
                            const int n16 = nlen >> 4;
                            const uint64_t dob = ab_mn_desc(do_u + (uint32_t)n0 * 128u + cb, 16384u), qbm = ab_mn_desc(q_u + (uint32_t)n0 * 128u + cb, 16384u);
                            for (int part = 0; part < 4; ++part) {
                                const int gb = part * n16 / 4, ge = (part + 1) * n16 / 4;
                                for (int gq = gb; gq < ge; ++gq) {
                                    const uint32_t acol = (uint32_t)(16 * gb + 8 * (gq - gb));
                                    const uint32_t acc = (qb > 0 || gq > 0) ? 1u : 0u;
                                    ab_umma_ts(tmem_base + (uint32_t)p.dv_col, tmem_base + acol, dob + (uint64_t)gq * 128u, idesc_ts, acc);
                                    ab_umma_ts(tmem_base + (uint32_t)p.dk_col, tmem_base + (uint32_t)p.dp_col + acol, qbm + (uint64_t)gq * 128u, idesc_ts, acc);
                                }
                            }
                            // dQ_qb (+)= dS K_j : M-tiles of 128 queries; A = shared dS^T tile (MN-major), K = the loaded keys of tile j
                            const uint64_t kbm = ab_mn_desc(k_u + (uint32_t)(128 * j) * 128u + cb, 16384u);
                            const int ks_n = kvalid >> 4;
                            for (int mt = 0; mt < p.qb_nmt[qb]; ++mt) {
                                const uint64_t ad = ab_mn_desc(dst_u + (uint32_t)p.mt_skb[p.qb_mt0[qb] + mt] * 16384u, 16384u);
                                const uint32_t dcol = tmem_base + (uint32_t)(p.dq_col + (p.qb_mt0[qb] + mt) * p.hp);
                                for (int ks = 0; ks < ks_n; ++ks)
                                    umma_bf16(dcol, ad + (uint64_t)ks * 128u, kbm + (uint64_t)ks * 128u, idesc_mm, (j > 0 || ks > 0) ? 1u : 0u);
                            }
                            if (qb == p.nqb - 1) umma_commit(grad_done);
                        }
                        __syncwarp();
                        AB_CLK(8 * mj + 3);
                        ++mj;
                    }
                }
            }
        }
        (void)ndrain_waits;
    } else {
        // ===== workers: thread = key row of the current 128-key tile; 4 warps per lane quarter split the query columns
        const int quarter = warp & 3, part = warp >> 2, row = quarter * 32 + lane;
        const int wtid = threadIdx.x;                                             // 0..511
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float c = p.sc * AB_LOG2E;
        const int sw = row & 7;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && warp == 0 && lane == 0;
        int wj = 0;
        uint32_t ph_sdp = 0, ph_gd = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it % p.nstage;
            const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
            const int b = item / p.NG, g = item - b * p.NG;
            const int nheads = min(p.G, p.H - g * p.G);
            const int row0 = b * p.Lp + p.halo;
            AB_CLK(500 + 8 * wj + 7);
            mbar_wait(&full[st], ph);
            // ---- per-item vectors: lse * log2(e) and D = rowsum(dO * O) per (head, query); bias-gradient partial sums
            {
                const uint8_t* do_s = smem + (size_t)st * p.stage_bytes + (size_t)3 * p.opnd_bytes;
                const int cpr = p.hp >> 3;                                        // 16-byte chunks per head row
                for (int i = wtid; i < nheads * p.Lk; i += 32 * AB_WORKER_WARPS) {
                    const int h = i / p.Lk, q = i - h * p.Lk, head = g * p.G + h;
                    float dsum = 0.f, l2 = 0.f;
                    if (q < p.L) {
                        const bf16* orow = p.o + (size_t)(row0 + q) * p.ldo + head * p.hp;
                        for (int cc = 0; cc < cpr; ++cc) {
                            const uint4 ov = *reinterpret_cast<const uint4*>(orow + cc * 8);
                            const int chunk = (h * p.hp * 2) / 16 + cc;
                            const uint4 gv = *reinterpret_cast<const uint4*>(do_s + (size_t)q * 128 + ((chunk ^ (q & 7)) << 4));
                            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
                            const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 x = __bfloat1622float2(oh[u]), y = __bfloat1622float2(gh[u]);
                                dsum = fmaf(x.x, y.x, dsum);
                                dsum = fmaf(x.y, y.y, dsum);
                            }
                        }
                        l2 = p.lse[((size_t)b * p.H + head) * p.L + q] * AB_LOG2E;
                    }
                    lse2[i] = l2;
                    Dv[i] = dsum;
                }
                for (int i = wtid; i < 192; i += 32 * AB_WORKER_WARPS) csum[i] = 0.f;
                named_bar_sync(1, 32 * AB_WORKER_WARPS);
            }
            for (int h = 0; h < nheads; ++h) {
                const int head = g * p.G + h;
                const float* l2h = lse2 + h * p.Lk;
                const float* Dh = Dv + h * p.Lk;
                float* cs_h = p.dbias ? csum + h * 3 * p.hp : nullptr;
                for (int j = 0; j < p.nkt; ++j) {
                    const int krow = 128 * j + row;
                    const bool kvalid = krow < p.L;                               // this thread's key exists
                    const bool qactive = 128 * j + quarter * 32 < p.Lk;           // warp-uniform: the quarter holds loaded key rows
                    for (int qb = 0; qb < p.nqb; ++qb) {
                        const int n0 = p.qb_n0[qb], nlen = p.qb_len[qb], n16 = nlen >> 4;
                        const int gb = part * n16 / 4, ge = (part + 1) * n16 / 4;
                        AB_CLK(500 + 8 * wj + 0);
                        mbar_wait(sdp_full, ph_sdp);
                        ph_sdp ^= 1u;
                        tc_fence_after();
                        AB_CLK(500 + 8 * wj + 1);
                        if (qactive) {
                            for (int gq = gb; gq < ge; ++gq) {
                                uint32_t rs[16], rp[16];
                                tmem_ld16(trow + (uint32_t)(16 * gq), rs);
                                tmem_ld16(trow + (uint32_t)(p.dp_col + 16 * gq), rp);
                                const int q0 = n0 + 16 * gq;
                                float lq[16], dq_[16];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float4 a = *reinterpret_cast<const float4*>(l2h + q0 + 4 * u), d4 = *reinterpret_cast<const float4*>(Dh + q0 + 4 * u);
                                    lq[4 * u] = a.x; lq[4 * u + 1] = a.y; lq[4 * u + 2] = a.z; lq[4 * u + 3] = a.w;
                                    dq_[4 * u] = d4.x; dq_[4 * u + 1] = d4.y; dq_[4 * u + 2] = d4.z; dq_[4 * u + 3] = d4.w;
                                }
                                tmem_ld_wait();
                                uint32_t pp[8], pd[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    float pv[2], dv[2];
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        const int jj = 2 * u + e;
                                        const bool ok = kvalid && (q0 + jj < p.L);
                                        const float pr = ab_exp2(fmaf(__uint_as_float(rs[jj]), c, -lq[jj]));
                                        pv[e] = ok ? pr : 0.f;
                                        dv[e] = ok ? pr * (__uint_as_float(rp[jj]) - dq_[jj]) : 0.f;
                                    }
                                    pp[u] = ab_pack(pv[0], pv[1]);
                                    pd[u] = ab_pack(dv[0], dv[1]);
                                }
                                const uint32_t pcol = (uint32_t)(16 * gb + 8 * (gq - gb));
                                ab_tmem_st8(trow + pcol, pp);
                                ab_tmem_st8(trow + (uint32_t)p.dp_col + pcol, pd);
                                // dS^T -> shared tile [key row][query], 64-query blocks of 16 KB, 128B swizzle
                                const uint32_t tb = dst_u + (uint32_t)(q0 >> 6) * 16384u + (uint32_t)row * 128u;
                                const int ch0 = (q0 & 63) >> 3;
                                ab_sts128(tb + (uint32_t)((ch0 ^ sw) << 4), pd[0], pd[1], pd[2], pd[3]);
                                ab_sts128(tb + (uint32_t)(((ch0 + 1) ^ sw) << 4), pd[4], pd[5], pd[6], pd[7]);
                            }
                            ab_tmem_st_wait();
                        }
                        tc_fence_before();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) ab_mbar_arrive(pds_ready);
                        AB_CLK(500 + 8 * wj + 2);
                    }
                    // ---- drain dV_j, dK_j (and dQ after the last key tile) once their products have completed
                    mbar_wait(grad_done, ph_gd);
                    ph_gd ^= 1u;
                    tc_fence_after();
                    AB_CLK(500 + 8 * wj + 3);
                    if (qactive) {
                        bf16* drow = p.dqkv + (size_t)(row0 + krow) * p.lddq + head * p.hp;
                        const int ngr = p.hp >> 4;                                // 16-column groups per accumulator
                        for (int gg = part; gg < 2 * ngr; gg += 4) {              // dV groups, then dK groups
                            const bool isk = gg >= ngr;
                            const int c16 = (isk ? gg - ngr : gg) * 16;
                            ab_drain16(trow + (uint32_t)((isk ? p.dk_col : p.dv_col) + c16), isk ? p.sc : 1.f, kvalid,
                                       drow + (isk ? p.HP : 2 * p.HP) + c16, cs_h ? cs_h + (isk ? p.hp : 2 * p.hp) + c16 : nullptr, lane);
                        }
                    }
                    if (j == p.nkt - 1) {
                        const int ngr = p.hp >> 4, nmt = p.qb_mt0[p.nqb - 1] + p.qb_nmt[p.nqb - 1];
                        for (int t = 0; t < nmt; ++t) {
                            const int qlo = 64 * p.mt_skb[t] + quarter * 32, qend = min(p.mt_q1[t], p.L);
                            if (qlo + 31 < p.mt_q0[t] || qlo >= qend) continue;                      // warp-uniform: no owned query in this quarter
                            const int qrow = qlo + lane;
                            const bool qvalid = qrow >= p.mt_q0[t] && qrow < qend;
                            bf16* drow = p.dqkv + (size_t)(row0 + qrow) * p.lddq + head * p.hp;
                            for (int gg = part; gg < ngr; gg += 4)
                                ab_drain16(trow + (uint32_t)(p.dq_col + t * p.hp + gg * 16), p.sc, qvalid, drow + gg * 16,
                                           cs_h ? cs_h + gg * 16 : nullptr, lane);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ab_mbar_arrive(drained);
                    AB_CLK(500 + 8 * wj + 4);
                    ++wj;
                }
            }
            // ---- item done: every product has completed (grad_done) and every accumulator has been read
            named_bar_sync(1, 32 * AB_WORKER_WARPS);
            if (p.dbias) {
                for (int i = wtid; i < nheads * 3 * p.hp; i += 32 * AB_WORKER_WARPS) {
                    const int h = i / (3 * p.hp), r = i - h * 3 * p.hp, w = r / p.hp, e = r - w * p.hp;
                    if (e < p.hd) atomicAdd(p.dbias + w * p.d + (g * p.G + h) * p.hd + e, csum[i]);
                }
            }
            if (wtid == 0) ab_mbar_arrive(&empty[st]);
            named_bar_sync(1, 32 * AB_WORKER_WARPS);              // csum is zeroed again at the start of the next item
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------- host side
struct AbGeom {
    int Lk, RB, nbox, nkt, G, NG, nstage, nqb, qb_n0[AB_MAX_QB], qb_len[AB_MAX_QB], qb_mt0[AB_MAX_QB], qb_nmt[AB_MAX_QB];
    int mt_skb[AB_MAX_MT], mt_q0[AB_MAX_MT], mt_q1[AB_MAX_MT];
    int dp_col, dv_col, dk_col, dq_col;
    uint32_t opnd_bytes, stage_bytes, dst_bytes;
    size_t smem;
};

static bool ab_geom(int L, int d, int H, int hp, AbGeom& g) {
    if (H <= 0 || d % H || L < 1) return false;
    const int hd = d / H;
    const int pad = hd <= 16 ? 16 : (hd <= 32 ? 32 : 64);
    if (hd > 64 || hp != pad) return false;
    g.Lk = (L + 15) & ~15;
    g.nbox = (g.Lk + 255) / 256;
    while ((g.Lk % g.nbox) || ((g.Lk / g.nbox) % 8)) ++g.nbox;
    g.RB = g.Lk / g.nbox;
    g.nkt = (L + 127) / 128;
    g.G = 64 / hp;
    g.NG = (H + g.G - 1) / g.G;
    g.opnd_bytes = ((uint32_t)g.Lk * 128u + 1023u) & ~1023u;
    g.stage_bytes = 4 * g.opnd_bytes;
    const int nkbq = max(2, (g.Lk + 63) / 64);                   // 64-query blocks of the dS^T tile (an M-tile reads two of them)
    g.dst_bytes = (uint32_t)nkbq * 16384u;
    // query blocks: 64-aligned starts, each at most 256 wide and a multiple of 16; the fewest blocks whose accumulators fit TMEM
    bool ok = false;
    for (int nblk = 1; nblk <= AB_MAX_QB && !ok; ++nblk) {
        const int base = nblk == 1 ? g.Lk : ((g.Lk + nblk - 1) / nblk + 63) / 64 * 64;
        if (nblk > 1 && (nblk - 1) * base >= g.Lk) continue;
        int nq = 0, nmt = 0;
        for (int i = 0; i < nblk; ++i) {
            g.qb_n0[i] = i * base;
            g.qb_len[i] = i == nblk - 1 ? g.Lk - i * base : base;
            g.qb_mt0[i] = nmt;
            g.qb_nmt[i] = (g.qb_len[i] + 127) / 128;
            for (int t = 0; t < g.qb_nmt[i] && nmt + t < AB_MAX_MT; ++t) {
                const int q0 = g.qb_n0[i] + 128 * t;
                g.mt_skb[nmt + t] = min(q0 / 64, nkbq - 2);
                g.mt_q0[nmt + t] = q0;
                g.mt_q1[nmt + t] = min(q0 + 128, g.qb_n0[i] + g.qb_len[i]);
            }
            nmt += g.qb_nmt[i];
            if (g.qb_len[i] > nq) nq = g.qb_len[i];
        }
        if (nq > 256 || nmt > AB_MAX_MT) continue;
        const int nqa = (nq + 31) & ~31;
        if (2 * nqa + 2 * hp + nmt * hp > 512) continue;
        g.nqb = nblk;
        g.dp_col = nqa; g.dv_col = 2 * nqa; g.dk_col = 2 * nqa + hp; g.dq_col = 2 * nqa + 2 * hp;
        ok = true;
    }
    if (!ok) return false;
    const size_t tail = (size_t)g.dst_bytes + 2 * (size_t)g.G * g.Lk * sizeof(float) + 192 * sizeof(float) + 128 + 1024;
    // every operand is read as whole 128-row tiles: rows past Lk fall into whatever follows the operand inside the allocation
    const size_t min_stage_span = (size_t)g.nkt * 16384 + 3 * (size_t)g.opnd_bytes;
    g.nstage = (2 * (size_t)g.stage_bytes + tail <= 227 * 1024) ? 2 : 1;
    g.smem = (size_t)g.nstage * g.stage_bytes + tail;
    if ((size_t)g.nstage * g.stage_bytes + g.dst_bytes < min_stage_span) return false;
    return g.smem <= 227 * 1024 && get_encode() != nullptr;
}

extern "C" int csi_attn_bwd_tc_ok(int L, int d, int H, int hp) {
    AbGeom g;
    return ab_geom(L, d, H, hp, g) ? 1 : 0;
}

static int g_ab_sms = 0;
static long long* g_ab_dbg = nullptr;
extern "C" int csi_set_attn_bwd_debug(long long* buf) { g_ab_dbg = buf; return CSI_OK; }

extern "C" int csi_attn_bwd_tc(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo, void* dqkv,
                               int lddqkv, const float* lse, int B, int L, int d, int H, int hp, int halo, float* dbias,
                               void* stream) {
    CSI_CHECK_ARG(qkv && o && dout && dqkv && lse, "null pointer");
    AbGeom g;
    CSI_CHECK_ARG(ab_geom(L, d, H, hp, g), "shape not eligible");
    CSI_CHECK_ARG(ld3 % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddqkv % 8 == 0 && ld3 >= 3 * H * hp && lddqkv >= 3 * H * hp &&
                  ldo >= H * hp && lddo >= H * hp, "head-padded leading dimensions expected");
    CSI_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout) |
                    reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0, "16-byte aligned buffers expected");
    if (B == 0) return CSI_OK;
    if (g_ab_sms == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_ab_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int Lp = L + 2 * halo;
    CUtensorMap tmQ, tmG;
    int rc = make_map(&tmQ, qkv, (long long)B * Lp, ld3, ld3, g.RB);
    if (rc) return rc;
    rc = make_map(&tmG, dout, (long long)B * Lp, lddo, lddo, g.RB);
    if (rc) return rc;
    AbParams p;
    p.B = B; p.L = L; p.Lk = g.Lk; p.Lp = Lp; p.H = H; p.hp = hp; p.hd = d / H; p.G = g.G; p.NG = g.NG; p.halo = halo;
    p.RB = g.RB; p.nbox = g.nbox; p.nkt = g.nkt; p.nqb = g.nqb;
    for (int i = 0; i < AB_MAX_QB; ++i) { p.qb_n0[i] = g.qb_n0[i]; p.qb_len[i] = g.qb_len[i]; p.qb_mt0[i] = g.qb_mt0[i]; p.qb_nmt[i] = g.qb_nmt[i]; }
    for (int i = 0; i < AB_MAX_MT; ++i) { p.mt_skb[i] = g.mt_skb[i]; p.mt_q0[i] = g.mt_q0[i]; p.mt_q1[i] = g.mt_q1[i]; }
    p.ldo = ldo; p.lddq = lddqkv; p.HP = H * hp; p.d = d;
    p.o = reinterpret_cast<const bf16*>(o); p.dqkv = reinterpret_cast<bf16*>(dqkv); p.lse = lse; p.dbias = dbias;
    p.sc = 1.0f / sqrtf((float)(d / H));
    p.opnd_bytes = g.opnd_bytes; p.stage_bytes = g.stage_bytes; p.dst_bytes = g.dst_bytes; p.nstage = g.nstage;
    p.dbg = g_ab_dbg;
    p.dp_col = g.dp_col; p.dv_col = g.dv_col; p.dk_col = g.dk_col; p.dq_col = g.dq_col;
    const int nitems = B * g.NG;
    const int grid = nitems < g_ab_sms ? nitems : g_ab_sms;
    CSI_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    attn_bwd_tc_kernel<<<grid, AB_THREADS, g.smem, ST(stream)>>>(tmQ, tmG, p);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
