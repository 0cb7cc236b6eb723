// EXPERIMENT, not part of libcsi_that.so (moved out of multi_modal_csi_b200/csrc at the end of round 2): build it by adding this file
// to SRCS in multi_modal_csi_b200/csrc/Makefile (it includes "tc_common.cuh" from there); scripts/test_attn_tc.py picks up
// csi_attn_fwd_tc2 when the library exports it.  Measured 71 us at B=256, L=150, d=270 against 51 us for the mma.sync forward.
// tcgen05 / TMEM attention forward, second generation (that.py:113-115,149; sequences of up to 160 tokens: the temporal stream).
//
// attention_tc.cu runs a (head, query tile) job end to end inside one team of warps: S MMA -> wait -> softmax -> PV MMA ->
// wait -> drain, so every job pays both MMA round trips and two TMEM passes (phase clocks: ~8 500 clk per job against
// ~1 300 clk of MUFU work).  Here the roles are split, FlashAttention-4 style, and two jobs are always in flight:
//
//   warp 9      TMA loader: Q | K | V boxes [Lk x 64 columns] of the next item (sample, head group) into a 2-stage ring
//   warp 8      MMA issuer (one elected lane).  Job n uses TMEM slot n & 1.  Issue order: S_n = Q K^T, then PV_{n-1}: the
//               score product of the next job is already in TMEM when its softmax warps become free, and the PV product of
//               a job runs while the other warpgroup does its softmax.  tcgen05.mma executes in issue order, so S_n may
//               overwrite the slot whose P_{n-2} the (earlier issued) PV_{n-2} reads.
//   warps 0-3   softmax warpgroup 0 (jobs 0, 2, 4, ...)      thread = query row = TMEM lane
//   warps 4-7   softmax warpgroup 1 (jobs 1, 3, 5, ...)
//               The WHOLE score row (<= 160 fp32) is read from TMEM once into registers (one wait), max / exp2 / sum run on
//               registers, P (bf16 pairs) goes back in place as the TMEM A operand of the PV product; later the warpgroup
//               drains O (scaled by 1/sum, bf16) and the log-sum-exp.
//   A ragged last query tile (L = 150: 22 rows) is shifted so that its rows land in lane quarter (job mod 4): the three
//   empty quarters skip the softmax, and over time every SM sub-partition gets the same exp load.
#include "tc_common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define A2_THREADS 320
#define A2_LOG2E 1.4426950408889634f

__device__ __forceinline__ uint64_t a2_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float a2_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t a2_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void a2_tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void a2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void a2_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct A2Params {
    int B, L, Lk, Lp, H, hp, hd, G, NG, halo, nqt;
    int ldo, HP;
    bf16* o;
    float* lse;
    float sc;
    uint32_t opnd_bytes;          // bytes of one [Lk x 64] operand box set, rounded up to 1024
    int s_pitch;                  // TMEM columns between the two S/P slots
    int o_col;                    // first TMEM column of the two O accumulators (hp columns each)
    long long* dbg;               // optional phase clocks of one CTA (scripts/test_attn_tc.py phases2), NULL in production
};
#define A2_CLK(slot) do { if (dbg_on && (slot) < 2000) p.dbg[(slot)] = clock64(); } while (0)

// job n of a CTA -> (head inside the group, query tile)
struct A2Job { int h, i; };

template <int NCH>
__global__ void __launch_bounds__(A2_THREADS, 1) attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, A2Params p) {
    extern __shared__ __align__(1024) uint8_t a2_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a2_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[2], empty[2], s_full[2], p_full[2], o_full[2], o_free[2];
    __shared__ uint32_t tmem_base_smem;
    const uint32_t stage_bytes = 3u * p.opnd_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = p.B * p.NG;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&o_full[s], 1); mbar_init(&o_free[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 9) {
        // ===== TMA loader
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it & 1;
            mbar_wait(&empty[st], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            if (lane == 0) {
                const int b = item / p.NG, g = item - b * p.NG;
                const int row0 = b * p.Lp + p.halo;
                mbar_expect_tx(&full[st], 3u * (uint32_t)p.Lk * 128u);
                uint8_t* dst = smem + (size_t)st * stage_bytes;
                for (int w = 0; w < 3; ++w)
                    tma_load_2d(&tmQKV, &full[st], dst + (size_t)w * p.opnd_bytes, w * p.HP + g * 64, row0);
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ===== MMA issuer (warp-uniform control flow, lane 0 issues)
        const int ksteps_s = p.hp >> 4;
        const uint32_t idesc_s = make_idesc(TC_BM, p.Lk);
        const uint32_t idesc_o = make_idesc(TC_BM, p.hp) | (1u << 16);               // B (= V) is MN-major
        uint32_t cnt[2] = {0, 0};                      // jobs issued per slot (barrier phases)
        int n = 0, it = 0;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0;
        // pending PV product of the previous job
        bool have_prev = false, prev_last = false;
        uint64_t prev_vb = 0;
        int prev_w = 0, prev_st = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int st = it & 1;
            const int g = item % p.NG;
            const int nheads = min(p.G, p.H - g * p.G);
            const int njobs = nheads * p.nqt;
            const uint32_t q_u = smem_u32(smem) + (uint32_t)st * stage_bytes, k_u = q_u + p.opnd_bytes, v_u = k_u + p.opnd_bytes;
            A2_CLK(1000 + 8 * n + 0);
            mbar_wait(&full[st], (uint32_t)(it >> 1) & 1u);
            for (int job = 0; job < njobs; ++job, ++n) {
                const int h = job / p.nqt, i = job - h * p.nqt;
                const int w = n & 1;
                A2_CLK(1000 + 8 * n + 1);
                const uint32_t cb = (uint32_t)(h * p.hp * 2);                        // head slice inside the 128-byte line
                // ragged last tile: start the 128-row window early so that its rows land in lane quarter (n mod 4)
                int q_row0 = i * 128;
                if (i == p.nqt - 1 && p.nqt > 1 && p.L - q_row0 <= 32) q_row0 -= 32 * (n & 3);
                if (lane == 0) {
                    tc_fence_after();
                    const uint64_t qa = make_kmajor_desc(q_u + (uint32_t)q_row0 * 128u + cb), ka = make_kmajor_desc(k_u + cb);
                    const uint32_t sacc = tmem_base + (uint32_t)(w * p.s_pitch);
                    for (int kk = 0; kk < ksteps_s; ++kk)
                        umma_bf16(sacc, qa + (uint64_t)(2 * kk), ka + (uint64_t)(2 * kk), idesc_s, kk ? 1u : 0u);
                    umma_commit(&s_full[w]);
                }
                __syncwarp();
                A2_CLK(1000 + 8 * n + 2);
                if (have_prev) {
                    // PV of the previous job: its P is complete, and the O accumulator of that slot has been drained (the
                    // (c-1)-th drain of the slot for its c-th job: parity (c-1) & 1 = cnt & 1, immediately true for c = 0)
                    mbar_wait(&p_full[prev_w], (cnt[prev_w] - 1u) & 1u);
                    A2_CLK(1000 + 8 * n + 3);
                    mbar_wait(&o_free[prev_w], (cnt[prev_w] & 1u));
                    A2_CLK(1000 + 8 * n + 4);
                    if (lane == 0) {
                        tc_fence_after();
                        const uint32_t pa = tmem_base + (uint32_t)(prev_w * p.s_pitch), oacc = tmem_base + (uint32_t)(p.o_col + prev_w * p.hp);
                        for (int ks = 0; ks < NCH; ++ks)
                            a2_umma_ts(oacc, pa + (uint32_t)(ks * 8), prev_vb + (uint64_t)ks * 128u, idesc_o, ks ? 1u : 0u);
                        umma_commit(&o_full[prev_w]);
                        if (prev_last) umma_commit(&empty[prev_st]);
                    }
                    __syncwarp();
                    A2_CLK(1000 + 8 * n + 5);
                }
                have_prev = true;
                prev_w = w; prev_st = st; prev_last = (job == njobs - 1);
                prev_vb = a2_mn_desc(v_u + cb, 16384u);
                ++cnt[w];
            }
        }
        if (have_prev) {
            mbar_wait(&p_full[prev_w], (cnt[prev_w] - 1u) & 1u);
            mbar_wait(&o_free[prev_w], (cnt[prev_w] & 1u));
            if (lane == 0) {
                tc_fence_after();
                const uint32_t pa = tmem_base + (uint32_t)(prev_w * p.s_pitch), oacc = tmem_base + (uint32_t)(p.o_col + prev_w * p.hp);
                for (int ks = 0; ks < NCH; ++ks)
                    a2_umma_ts(oacc, pa + (uint32_t)(ks * 8), prev_vb + (uint64_t)ks * 128u, idesc_o, ks ? 1u : 0u);
                umma_commit(&o_full[prev_w]);
                if (prev_last) umma_commit(&empty[prev_st]);
            }
            __syncwarp();
        }
    } else {
        // ===== softmax warpgroups
        const int w = warp >> 2, quarter = warp & 3;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t srow = tmem_base + (uint32_t)(w * p.s_pitch) + lane_off;
        const uint32_t orow = tmem_base + (uint32_t)(p.o_col + w * p.hp) + lane_off;
        const float c = p.sc * A2_LOG2E;
        uint32_t cnt = 0;                              // jobs of this warpgroup so far
        int n = 0;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && quarter == 0 && lane == 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int b = item / p.NG, g = item - b * p.NG;
            const int nheads = min(p.G, p.H - g * p.G);
            const int njobs = nheads * p.nqt;
            const int row0 = b * p.Lp + p.halo;
            for (int job = 0; job < njobs; ++job, ++n) {
                if ((n & 1) != w) continue;
                const int h = job / p.nqt, i = job - h * p.nqt;
                int q_row0 = i * 128;
                if (i == p.nqt - 1 && p.nqt > 1 && p.L - q_row0 <= 32) q_row0 -= 32 * (n & 3);
                const int qrow = q_row0 + quarter * 32 + lane;                       // query row of this thread
                // rows below the tile's own range were recomputed by the shifted window: they belong to the previous tile
                const bool qvalid_w = (q_row0 + quarter * 32 + 31 >= i * 128) && (q_row0 + quarter * 32 < p.L);   // warp-uniform
                const bool rvalid = qrow >= i * 128 && qrow < p.L;
                A2_CLK(8 * n + 0);
                mbar_wait(&s_full[w], cnt & 1u);
                tc_fence_after();
                A2_CLK(8 * n + 1);
                float sum = 1.f, m = 0.f;
                if (qvalid_w) {
                    uint32_t r[NCH * 16];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) tmem_ld16(srow + (uint32_t)(ch * 16), r + ch * 16);
                    tmem_ld_wait();
                    m = -INFINITY;
#pragma unroll
                    for (int j = 0; j < NCH * 16; ++j)                      // only the last 16-key group can hold padding keys
                        if (j < (NCH - 1) * 16 || j < p.L) m = fmaxf(m, __uint_as_float(r[j]));
                    const float mc = m * c;
                    sum = 0.f;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int j0 = ch * 16 + 2 * j;
                            float e0 = a2_exp2(fmaf(__uint_as_float(r[j0]), c, -mc)), e1 = a2_exp2(fmaf(__uint_as_float(r[j0 + 1]), c, -mc));
                            if (ch == NCH - 1) {                                      // only the last 16-key group can hold padding keys
                                if (j0 >= p.L) e0 = 0.f;
                                if (j0 + 1 >= p.L) e1 = 0.f;
                            }
                            sum += e0 + e1;
                            pk[j] = a2_pack(e0, e1);
                        }
                        a2_tmem_st8(srow + (uint32_t)(ch * 8), pk);
                    }
                    a2_tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[w]);
                A2_CLK(8 * n + 2);
                // ---- O / rowsum -> bf16 -> global; log-sum-exp
                mbar_wait(&o_full[w], cnt & 1u);
                tc_fence_after();
                A2_CLK(8 * n + 3);
                if (qvalid_w) {
                    const float inv = 1.f / sum;
                    const int head = g * p.G + h;
                    for (int oc = 0; oc < p.hp; oc += 16) {
                        uint32_t ro[16];
                        tmem_ld16(orow + (uint32_t)oc, ro);
                        tmem_ld_wait();
                        if (rvalid) {
                            bf16* dst = p.o + (size_t)(row0 + qrow) * p.ldo + head * p.hp + oc;
#pragma unroll
                            for (int hlf = 0; hlf < 2; ++hlf) {
                                uint4 u;
                                u.x = a2_pack(__uint_as_float(ro[hlf * 8 + 0]) * inv, __uint_as_float(ro[hlf * 8 + 1]) * inv);
                                u.y = a2_pack(__uint_as_float(ro[hlf * 8 + 2]) * inv, __uint_as_float(ro[hlf * 8 + 3]) * inv);
                                u.z = a2_pack(__uint_as_float(ro[hlf * 8 + 4]) * inv, __uint_as_float(ro[hlf * 8 + 5]) * inv);
                                u.w = a2_pack(__uint_as_float(ro[hlf * 8 + 6]) * inv, __uint_as_float(ro[hlf * 8 + 7]) * inv);
                                *reinterpret_cast<uint4*>(dst + hlf * 8) = u;
                            }
                        }
                    }
                    if (rvalid) p.lse[((size_t)b * p.H + head) * p.L + qrow] = m * p.sc + __logf(sum);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free[w]);
                A2_CLK(8 * n + 4);
                ++cnt;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------- host side
static int a2_pad_hd(int hd) { return hd <= 16 ? 16 : (hd <= 32 ? 32 : 64); }

extern "C" int csi_attn_tc2_ok(int L, int d, int H, int hp) {
    if (H <= 0 || d % H || L < 1) return 0;
    const int hd = d / H;
    if (hd > 64 || hp != a2_pad_hd(hd)) return 0;
    const int Lk = (L + 15) & ~15;
    if (Lk > 160) return 0;                                   // the score row lives in registers (<= 160 fp32)
    // a ragged last tile is shifted back by up to 96 rows: it must be the only partial tile (L <= 160 guarantees nqt <= 2)
    return get_encode() != nullptr;
}

static int g_a2_sms = 0;
static long long* g_a2_dbg = nullptr;
extern "C" int csi_set_attn2_debug(long long* buf) { g_a2_dbg = buf; return CSI_OK; }

extern "C" int csi_attn_fwd_tc2(const void* qkv, int ld3, void* o, int ldo, float* lse, int B, int L, int d, int H, int hp,
                                int halo, void* stream) {
    CSI_CHECK_ARG(qkv && o && lse, "null pointer");
    CSI_CHECK_ARG(csi_attn_tc2_ok(L, d, H, hp), "shape not eligible");
    CSI_CHECK_ARG(ld3 % 8 == 0 && ldo % 8 == 0 && ld3 >= 3 * H * hp && ldo >= H * hp, "head-padded leading dimensions expected");
    CSI_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "16-byte aligned buffers expected");
    if (B == 0) return CSI_OK;
    if (g_a2_sms == 0) {
        int dev = 0;
        CSI_CUDA(cudaGetDevice(&dev));
        CSI_CUDA(cudaDeviceGetAttribute(&g_a2_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int Lk = (L + 15) & ~15, Lp = L + 2 * halo;
    CUtensorMap tm;
    int rc = make_map(&tm, qkv, (long long)B * Lp, ld3, ld3, Lk);
    if (rc) return rc;
    A2Params p;
    p.B = B; p.L = L; p.Lk = Lk; p.Lp = Lp; p.H = H; p.hp = hp; p.hd = d / H; p.G = 64 / hp; p.NG = (H + p.G - 1) / p.G; p.halo = halo;
    p.nqt = (L + 127) / 128;
    p.ldo = ldo; p.HP = H * hp;
    p.o = reinterpret_cast<bf16*>(o); p.lse = lse; p.sc = 1.0f / sqrtf((float)(d / H));
    p.opnd_bytes = ((uint32_t)Lk * 128u + 1023u) & ~1023u;
    p.s_pitch = (Lk + 31) & ~31;
    p.o_col = 2 * p.s_pitch;
    p.dbg = g_a2_dbg;
    // Q is read as whole 128-row tiles: the rows past Lk of the last operand of a stage fall into the next stage / the slack
    const size_t smem = 2 * 3 * (size_t)p.opnd_bytes + 16384 + 1024;
    const int nitems = B * p.NG;
    const int grid = nitems < g_a2_sms ? nitems : g_a2_sms;
#define A2_GO(N)                                                                                                        \
    case N:                                                                                                             \
        CSI_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        attn_fwd_tc2_kernel<N><<<grid, A2_THREADS, smem, ST(stream)>>>(tm, p);                                          \
        break;
    switch (Lk >> 4) {
        A2_GO(1) A2_GO(2) A2_GO(3) A2_GO(4) A2_GO(5) A2_GO(6) A2_GO(7) A2_GO(8) A2_GO(9) A2_GO(10)
        default: csi_set_error("csi_attn_fwd_tc2: unsupported length"); return CSI_ERR_ARG;
    }
#undef A2_GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
