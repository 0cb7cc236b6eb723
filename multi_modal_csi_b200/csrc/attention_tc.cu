// tcgen05 / TMEM attention core for bf16 activations (that.py:113-115,149: nn.MultiheadAttention, weights discarded).
//
// One CTA owns one sample and one HEAD GROUP: the G = 64/hp heads whose q (k, v) columns share one 128-byte line of the
// head-padded qkv token buffer.  Q, K and V of the group arrive as three TMA boxes [Lk rows x 64 columns] in the 128B
// swizzle; a head is a 2*hp-byte column slice of those boxes, addressed by advancing the UMMA descriptors inside the
// swizzle atom (the swizzle is a function of the absolute shared-memory address).
//
//   forward, per (head, 128-query tile):
//     S = Q K^T          tcgen05.mma  M=128, N=Lk (two instructions when Lk > 256), K = hp       -> TMEM columns [0, Lk)
//     softmax            thread = query row = TMEM lane: tcgen05.ld, row max, exp2, row sum; P (bf16) is written to shared
//                        memory as a K-major 128B-swizzled operand (64-key blocks)
//     O = P V            tcgen05.mma  M=128, N=hp, K = Lk; V is read in place as an MN-major operand   -> TMEM [Lk, Lk+hp)
//     O / rowsum -> bf16 -> global, lse = max*scale + log(rowsum)
//   The S product of the next (head, tile) is issued together with the PV product of the current one, so the tensor pipe
//   works while the threads drain O; co-resident CTAs (two per SM when Lk <= 160) overlap their softmax and MMA phases.
//
//   backward, per (head, 128-query tile i, key block j):   (lse comes from the forward pass: no row max needed)
//     S = Q_i K_j^T, dP = dO_i V_j^T          -> TMEM
//     P = exp2(S*c - lse), dS = P * (dP - D_i) -> bf16 in shared memory, both as [query][key] tiles
//     dV_j += P^T dO_i, dK_j += dS^T Q_i      (P / dS read as MN-major A operands, dO_i / Q_i as MN-major B operands)
//     dQ_i += dS K_j                           (dS K-major A, K_j MN-major B)
//   dQ/dK/dV accumulate in TMEM across the blocks and are scaled, rounded to bf16 and stored once; the in_proj bias
//   gradient (column sums of the stored dq|dk|dv) is reduced per CTA and added with one atomic per channel.
#include "tc_common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define AT_THREADS 128
#define AT_LOG2E 1.4426950408889634f

// MN-major 128B-swizzled operand (rows = K index, 64 MN elements per 128-byte row); lbo = distance between 64-element
// MN chunks (unused when the MN extent is <= 64)
__device__ __forceinline__ uint64_t at_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float at_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void at_sts128(uint32_t saddr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t at_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// 8 packed 32-bit columns (16 bf16 values) of this thread's TMEM lane
__device__ __forceinline__ void at_tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void at_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]   (A: bf16 pairs packed in 32-bit columns, lane = row, K-major)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define AT_MAX_TEAMS 3
#define AT_MAX_SPLIT 4
#define AT_MAX_WARPS 20            // 16 softmax warps + the loader warp, rounded up
struct AtParams {
    int B, L, Lk, Lp, H, hp, hd, G, NG, halo;
    int RB, nbox;                 // TMA box rows and boxes per operand (Lk = RB * nbox)
    int nqt;                      // 128-query tiles per head
    int ldo, HP;                  // HP = H * hp
    bf16* o;
    float* lse;
    float sc;                     // 1 / sqrt(hd)
    uint32_t opnd_bytes;          // bytes of one [Lk x 64] operand, rounded up to 1024
    uint32_t smem_tiles;          // bytes of the operand ring (the exchange arrays and barriers follow it)
    int nstage;                   // Q/K/V ring depth (items in flight)
    int nteam, nsplit;            // softmax teams per CTA; warps per TMEM lane quarter inside a team (key columns split)
    int slot_cols, p_col, o_col;  // TMEM columns of one team slot; offsets of P and O inside it (S sits at 0)
    long long* dbg;               // optional phase clocks of one CTA (scripts/test_attn_tc.py), NULL in production
};

// 16 score columns of one row: running max (MASK: only columns < nvalid count)
template <bool MASK> __device__ __forceinline__ float at_max16(const uint32_t* r, float m, int nvalid) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (!MASK || j < nvalid) m = fmaxf(m, __uint_as_float(r[j]));
    return m;
}
// 16 score columns -> probabilities: bf16 pairs into 8 TMEM columns (the A operand of the PV product); returns their fp32 sum
template <bool MASK> __device__ __forceinline__ float at_exp16(const uint32_t* r, float c, float mc, int nvalid, uint32_t taddr) {
    float sum = 0.f;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float e0 = at_exp2(fmaf(__uint_as_float(r[2 * j]), c, -mc)), e1 = at_exp2(fmaf(__uint_as_float(r[2 * j + 1]), c, -mc));
        if (MASK && 2 * j >= nvalid) e0 = 0.f;
        if (MASK && 2 * j + 1 >= nvalid) e1 = 0.f;
        sum += e0 + e1;
        pk[j] = at_pack(e0, e1);
    }
    at_tmem_st8(taddr, pk);
    return sum;
}

// ------------------------------------------------------------------------------------------------------- forward
// Persistent: CTA c walks the items (sample, head group) c, c + grid, ...  The last warp is the TMA loader (Q/K/V boxes of
// the next item land while the current one is processed); the other warps form `nteam` teams of 4*nsplit warps.  A team
// owns one TMEM slot and runs the (head, query tile) jobs dealt to it end to end -- its first lane also issues the team's
// MMAs -- so the teams drift apart and one team's MMA / barrier latencies are covered by the others' softmax.
#define AT_CLK(slot) do { if (dbg_on) p.dbg[(slot)] = clock64(); } while (0)
__global__ void __launch_bounds__(32 * AT_MAX_WARPS, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, AtParams p) {
    extern __shared__ __align__(1024) uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_bytes = 3u * p.opnd_bytes;
    // row max / row sum exchange between the warps of a lane quarter: [part][128] each (only with nsplit > 1, i.e. one team)
    float* xmax = reinterpret_cast<float*>(smem + (size_t)p.smem_tiles);
    float* xsum = xmax + AT_MAX_SPLIT * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + AT_MAX_SPLIT * 128);
    uint64_t* full = bars;                       // [2]
    uint64_t* empty = bars + 2;                  // [2]
    uint64_t* bar_s = bars + 4;                  // [AT_MAX_TEAMS]
    uint64_t* bar_o = bars + 4 + AT_MAX_TEAMS;   // [AT_MAX_TEAMS]
    uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 4 + 2 * AT_MAX_TEAMS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team_warps = 4 * p.nsplit, nsoft = p.nteam * team_warps;
    const int nitems = p.B * p.NG;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)p.nteam); }
        for (int t = 0; t < AT_MAX_TEAMS; ++t) { mbar_init(&bar_s[t], 1); mbar_init(&bar_o[t], 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == nsoft) {
        // ===== TMA loader (warp-uniform loop, lane 0 issues)
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it % p.nstage;
            const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
            mbar_wait(&empty[st], ph ^ 1u);
            if (lane == 0) {
                const int b = item / p.NG, g = item - b * p.NG;
                const int row0 = b * p.Lp + p.halo;
                mbar_expect_tx(&full[st], 3u * (uint32_t)p.Lk * 128u);
                uint8_t* dst = smem + (size_t)st * stage_bytes;
                for (int w = 0; w < 3; ++w)
                    for (int bx = 0; bx < p.nbox; ++bx)
                        tma_load_2d(&tmQKV, &full[st], dst + (size_t)w * p.opnd_bytes + (size_t)bx * p.RB * 128, w * p.HP + g * 64,
                                    row0 + bx * p.RB);
            }
            __syncwarp();
        }
    } else if (warp < nsoft) {
        // ===== softmax teams
        const int team = warp / team_warps, wt = warp - team * team_warps;
        const int quarter = warp & 3, part = wt >> 2, row = quarter * 32 + lane;      // row of the 128-query tile = TMEM lane
        const bool leader = wt == 0 && lane == 0;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && team == 0 && wt == team_warps - 1 && lane == 0;
        const int tthreads = 32 * team_warps;                                        // named barrier 1 + team
        const int qbar = 1 + AT_MAX_TEAMS + quarter;                                 // named barrier of a lane quarter (nteam == 1 when nsplit > 1)
        const uint32_t slot = tmem_base + (uint32_t)(team * p.slot_cols);
        const uint32_t trow = slot + ((uint32_t)(quarter * 32) << 16);
        const float c = p.sc * AT_LOG2E;
        const int ksteps_s = p.hp >> 4, ksteps_o = p.Lk >> 4;
        const uint32_t idesc_o = make_idesc(TC_BM, p.hp) | (1u << 16);            // B (= V) is MN-major
        const int s_n0 = p.Lk > 256 ? 144 : p.Lk, s_n1 = p.Lk - s_n0;            // S = one or two instructions per K step
        const uint32_t idesc_s0 = make_idesc(TC_BM, s_n0), idesc_s1 = make_idesc(TC_BM, s_n1 > 0 ? s_n1 : 16);
        // this warp's share of the key columns, in 16-column groups
        const int n16 = p.Lk >> 4;
        const int cbeg = (part * n16 / p.nsplit) << 4, cend = ((part + 1) * n16 / p.nsplit) << 4;
        uint32_t ph_s = 0, ph_o = 0;
        int it = 0, jc = 0;
        AT_CLK(0);
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it % p.nstage;
            const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
            const int b = item / p.NG, g = item - b * p.NG;
            const int nheads = min(p.G, p.H - g * p.G);
            const int row0 = b * p.Lp + p.halo;
            const int njobs = nheads * p.nqt;
            const uint32_t q_u = smem_u32(smem) + (uint32_t)st * stage_bytes, k_u = q_u + p.opnd_bytes, v_u = k_u + p.opnd_bytes;
            const uint64_t q_desc0 = make_kmajor_desc(q_u), k_desc0 = make_kmajor_desc(k_u), v_desc0 = at_mn_desc(v_u, 16384u);
            mbar_wait(&full[st], ph);
            // jobs of an item are dealt round-robin, rotated by the item counter so that every team sees the same mix
            for (int job = (team + p.nteam - it % p.nteam) % p.nteam; job < njobs; job += p.nteam) {
                const int h = job / p.nqt, i = job - h * p.nqt;
                const uint64_t cb = (uint64_t)(h * p.hp * 2) >> 4;                // head slice inside the 128-byte line, 16-byte units
                AT_CLK(16 + 8 * jc);
                if (leader) {
                    // S = Q_i K^T into the slot (the previous job's O has been drained: team barrier at the end of the loop)
                    tc_fence_after();
                    const uint64_t qa = q_desc0 + (uint64_t)i * 1024u + cb, ka = k_desc0 + cb;
#pragma unroll 4
                    for (int kk = 0; kk < ksteps_s; ++kk)
                        umma_bf16(slot, qa + (uint64_t)(2 * kk), ka + (uint64_t)(2 * kk), idesc_s0, kk ? 1u : 0u);
                    if (s_n1 > 0) {
                        const uint64_t kb2 = ka + (uint64_t)(s_n0 * 8);           // s_n0 rows of 128 bytes
#pragma unroll 4
                        for (int kk = 0; kk < ksteps_s; ++kk)
                            umma_bf16(slot + (uint32_t)s_n0, qa + (uint64_t)(2 * kk), kb2 + (uint64_t)(2 * kk), idesc_s1, kk ? 1u : 0u);
                    }
                    umma_commit(&bar_s[team]);
                }
                mbar_wait(&bar_s[team], ph_s);
                ph_s ^= 1u;
                tc_fence_after();
                AT_CLK(17 + 8 * jc);
                // quarters of a ragged last tile that hold no query row skip the softmax (their P rows stay stale: unused lanes)
                const bool qvalid = i * 128 + quarter * 32 < p.L;                // warp-uniform, equal for the warps of a quarter
                float m = -INFINITY, sum = 0.f;
                if (qvalid) {
                    // ---- pass 1: row max over the valid keys of this warp's columns
                    for (int c0 = cbeg; c0 < cend; c0 += 32) {
                        uint32_t r[32];
                        const bool two = c0 + 16 < cend;
                        tmem_ld16(trow + (uint32_t)c0, r);
                        if (two) tmem_ld16(trow + (uint32_t)(c0 + 16), r + 16);
                        tmem_ld_wait();
                        m = (c0 + 16 <= p.L) ? at_max16<false>(r, m, 16) : at_max16<true>(r, m, p.L - c0);
                        if (two) m = (c0 + 32 <= p.L) ? at_max16<false>(r + 16, m, 16) : at_max16<true>(r + 16, m, p.L - c0 - 16);
                    }
                    if (p.nsplit > 1) {
                        xmax[part * 128 + row] = m;
                        named_bar_sync(qbar, 32 * p.nsplit);
#pragma unroll
                        for (int s2 = 0; s2 < AT_MAX_SPLIT; ++s2)
                            if (s2 < p.nsplit) m = fmaxf(m, xmax[s2 * 128 + row]);
                    }
                    AT_CLK(18 + 8 * jc);
                    // ---- pass 2: p = exp2(s*c - m*c), row sum, P (bf16 pairs) -> TMEM
                    const float mc = m * c;
                    const uint32_t prow = trow + (uint32_t)p.p_col;
                    for (int c0 = cbeg; c0 < cend; c0 += 32) {
                        uint32_t r[32];
                        const bool two = c0 + 16 < cend;
                        tmem_ld16(trow + (uint32_t)c0, r);
                        if (two) tmem_ld16(trow + (uint32_t)(c0 + 16), r + 16);
                        tmem_ld_wait();
                        sum += (c0 + 16 <= p.L) ? at_exp16<false>(r, c, mc, 16, prow + (uint32_t)(c0 >> 1))
                                                : at_exp16<true>(r, c, mc, p.L - c0, prow + (uint32_t)(c0 >> 1));
                        if (two) sum += (c0 + 32 <= p.L) ? at_exp16<false>(r + 16, c, mc, 16, prow + (uint32_t)((c0 + 16) >> 1))
                                                         : at_exp16<true>(r + 16, c, mc, p.L - c0 - 16, prow + (uint32_t)((c0 + 16) >> 1));
                    }
                    at_tmem_st_wait();
                    if (p.nsplit > 1) xsum[part * 128 + row] = sum;
                }
                AT_CLK(19 + 8 * jc);
                tc_fence_before();
                named_bar_sync(1 + team, tthreads);
                if (leader) {
                    // O = P V : A = P from TMEM (8 columns per K step of 16 keys), B = V in place (MN-major)
                    tc_fence_after();
                    const uint64_t vb = v_desc0 + cb;
                    const uint32_t pa = slot + (uint32_t)p.p_col, tacc = slot + (uint32_t)p.o_col;
#pragma unroll 4
                    for (int ks = 0; ks < ksteps_o; ++ks)
                        umma_ts(tacc, pa + (uint32_t)(ks * 8), vb + (uint64_t)ks * 128u, idesc_o, ks ? 1u : 0u);
                    umma_commit(&bar_o[team]);
                }
                if (p.nsplit > 1 && qvalid) {
                    sum = 0.f;
#pragma unroll
                    for (int s2 = 0; s2 < AT_MAX_SPLIT; ++s2)
                        if (s2 < p.nsplit) sum += xsum[s2 * 128 + row];
                }
                mbar_wait(&bar_o[team], ph_o);
                ph_o ^= 1u;
                tc_fence_after();
                AT_CLK(20 + 8 * jc);
                // ---- O / rowsum -> bf16 -> global (16 output columns per warp of the quarter); lse
                const int qrow = i * 128 + row, head = g * p.G + h;
                if (qvalid) {
                    const float inv = 1.f / sum;
                    for (int oc = part * 16; oc < p.hp; oc += 16 * p.nsplit) {         // warp-uniform
                        uint32_t ro[16];
                        tmem_ld16(trow + (uint32_t)(p.o_col + oc), ro);
                        tmem_ld_wait();
                        if (qrow < p.L) {
                            bf16* orow = p.o + (size_t)(row0 + qrow) * p.ldo + head * p.hp + oc;
#pragma unroll
                            for (int ch = 0; ch < 2; ++ch) {
                                uint4 u;
                                u.x = at_pack(__uint_as_float(ro[ch * 8 + 0]) * inv, __uint_as_float(ro[ch * 8 + 1]) * inv);
                                u.y = at_pack(__uint_as_float(ro[ch * 8 + 2]) * inv, __uint_as_float(ro[ch * 8 + 3]) * inv);
                                u.z = at_pack(__uint_as_float(ro[ch * 8 + 4]) * inv, __uint_as_float(ro[ch * 8 + 5]) * inv);
                                u.w = at_pack(__uint_as_float(ro[ch * 8 + 6]) * inv, __uint_as_float(ro[ch * 8 + 7]) * inv);
                                *reinterpret_cast<uint4*>(orow + ch * 8) = u;
                            }
                        }
                    }
                    if (part == 0 && qrow < p.L) p.lse[((size_t)b * p.H + head) * p.L + qrow] = m * p.sc + __logf(sum);
                }
                tc_fence_before();
                named_bar_sync(1 + team, tthreads);          // O (and the exchanged sums) are drained before the slot is reused
                AT_CLK(21 + 8 * jc);
                if (jc < 50) ++jc;
            }
            // this team has issued its last read of the stage (its PV products have completed): hand it back to the loader
            if (leader) at_mbar_arrive(&empty[st]);
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
static int at_pad_hd(int hd) { return hd <= 16 ? 16 : (hd <= 32 ? 32 : 64); }

struct AtGeom { int Lk, RB, nbox, nqt, G, NG, nstage, nteam, nsplit, slot_cols, p_col, o_col; uint32_t opnd_bytes; size_t smem, smem_tiles; };

static bool at_geom_fwd(int L, int d, int H, int hp, AtGeom& g) {
    if (H <= 0 || d % H || L < 1) return false;
    const int hd = d / H;
    if (hd > 64 || hp != at_pad_hd(hd)) return false;
    g.Lk = (L + 15) & ~15;
    if (g.Lk > 400) return false;
    g.nbox = (g.Lk + 255) / 256;
    while ((g.Lk % g.nbox) || ((g.Lk / g.nbox) % 8)) ++g.nbox;    // equal boxes of a multiple of 8 rows
    g.RB = g.Lk / g.nbox;
    g.nqt = (L + 127) / 128;
    g.G = 64 / hp;
    g.NG = (H + g.G - 1) / g.G;
    g.opnd_bytes = ((uint32_t)g.Lk * 128u + 1023u) & ~1023u;
    // TMEM slot of a team.  One warp per lane quarter: P overwrites S in place (the thread that writes a P column has
    // already consumed the S columns under it) and O lands in the dead upper part of S.  With the key columns split over
    // several warps of a quarter, P gets its own columns behind S.
    const int in_place_o = (g.Lk / 2 + 31) & ~31;
    const int slot1 = (max(g.Lk, in_place_o + hp) + 31) & ~31;
    if (g.Lk <= 192 && slot1 <= 512) {
        g.nsplit = 1; g.p_col = 0; g.o_col = in_place_o; g.slot_cols = slot1;
        g.nteam = 512 / slot1;
        if (g.nteam > AT_MAX_TEAMS) g.nteam = AT_MAX_TEAMS;
    } else {
        g.nsplit = 4; g.nteam = 1;
        g.p_col = g.Lk; g.o_col = (g.Lk + g.Lk / 2 + 31) & ~31; g.slot_cols = g.o_col + hp;
        if (g.slot_cols > 512) return false;
    }
    const size_t tail = 2 * 4 * 128 * sizeof(float) + 128 + 1024;                // xmax/xsum, barriers, alignment slack
    const size_t stage = 3 * (size_t)g.opnd_bytes;
    const size_t min_tile = (size_t)g.nqt * 16384;                              // query tiles are read as whole 128-row tiles
    g.nstage = (2 * stage + tail <= 227 * 1024) ? 2 : 1;
    g.smem_tiles = g.nstage * stage < min_tile ? min_tile : g.nstage * stage;
    g.smem = g.smem_tiles + tail;
    return g.smem <= 227 * 1024 && get_encode() != nullptr;
}

static long long* g_at_dbg = nullptr;
extern "C" int csi_set_attn_debug(long long* buf) { g_at_dbg = buf; return CSI_OK; }
static int g_at_teams = 0;          // > 0: force the number of teams (A/B runs)
extern "C" int csi_set_attn_teams(int n) { g_at_teams = n; return CSI_OK; }
static int g_at_sms = 0;

extern "C" int csi_attn_tc_ok(int L, int d, int H, int hp) {
    AtGeom g;
    return at_geom_fwd(L, d, H, hp, g) ? 1 : 0;
}

extern "C" int csi_attn_fwd_tc(const void* qkv, int ld3, void* o, int ldo, float* lse, int B, int L, int d, int H, int hp,
                               int halo, void* stream) {
    CSI_CHECK_ARG(qkv && o && lse, "null pointer");
    AtGeom g;
    CSI_CHECK_ARG(at_geom_fwd(L, d, H, hp, g), "shape not eligible");
    CSI_CHECK_ARG(ld3 % 8 == 0 && ldo % 8 == 0 && ld3 >= 3 * H * hp && ldo >= H * hp, "head-padded leading dimensions expected");
    CSI_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "16-byte aligned buffers expected");
    if (B == 0) return CSI_OK;
    if (g_at_sms == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_at_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int Lp = L + 2 * halo;
    CUtensorMap tm;
    int rc = make_map(&tm, qkv, (long long)B * Lp, ld3, ld3, g.RB);
    if (rc) return rc;
    AtParams p;
    p.B = B; p.L = L; p.Lk = g.Lk; p.Lp = Lp; p.H = H; p.hp = hp; p.hd = d / H; p.G = g.G; p.NG = g.NG; p.halo = halo;
    p.RB = g.RB; p.nbox = g.nbox; p.nqt = g.nqt; p.ldo = ldo; p.HP = H * hp;
    p.o = reinterpret_cast<bf16*>(o); p.lse = lse; p.sc = 1.0f / sqrtf((float)(d / H));
    p.opnd_bytes = g.opnd_bytes; p.smem_tiles = (uint32_t)g.smem_tiles; p.nstage = g.nstage; p.nteam = g.nteam; p.nsplit = g.nsplit;
    if (g_at_teams > 0 && g_at_teams < p.nteam) p.nteam = g_at_teams;
    p.slot_cols = g.slot_cols; p.p_col = g.p_col; p.o_col = g.o_col; p.dbg = g_at_dbg;
    const int nitems = B * g.NG;
    const int grid = nitems < g_at_sms ? nitems : g_at_sms;
    const int threads = 32 * (p.nteam * 4 * p.nsplit + 1);
    CSI_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    attn_fwd_tc_kernel<<<grid, threads, g.smem, ST(stream)>>>(tm, p);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
