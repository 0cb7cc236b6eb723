// Tensor-core attention core for bf16 activations (that.py:149, nn.MultiheadAttention with need_weights discarded):
// one CTA per (sample, head); the head's Q, K, V (and dO in backward) live in shared memory, scores/probabilities
// only ever exist in registers.  Matrix products use mma.sync.m16n8k16 (bf16 in, fp32 accumulate) fed by ldmatrix.
// Heads are stored with a pitch of HDP = 16/32/64 elements (head dims 15/27/54 zero-padded by the in-projection's
// weight layout), so every head row is a whole number of aligned 16-byte chunks: tiles move with 128-bit accesses.
//   forward : flash-style online softmax over 64-key blocks, one 16-query tile per warp iteration
//   backward: pass A (per 16-query tile)  S, dP -> dS -> dQ = dS K
//             pass B (per 16-key tile)    S^T, dP^T -> dV = P^T dO, dK = dS^T Q     (no atomics, no cross-warp sums)
//   dK/dV tiles are staged over the K/V rows they belong to once every warp has finished pass A (mbarrier split barrier).
#include "common.cuh"
#include <stdlib.h>

#define ST(s) ((cudaStream_t)(s))
#define AM_MAX_WARPS 5
#ifndef AM_DIRECT_STORE
#define AM_DIRECT_STORE 1
#endif
// occupancy: registers are the limit (3 CTAs/SM at ~103 regs).  CTAs have at most 5 warps; the (160 threads, 4 CTAs) bound
// = 96 registers gives 20 resident warps per SM; the 64-wide heads need more accumulators and keep the loose bound
template <int HDP> struct am_bounds { static constexpr int min_ctas = HDP <= 32 ? 4 : 2; };
// backward keeps twice the accumulators: 32-wide heads spill under the 96-register bound, 3 CTAs/SM (128 registers) is faster
template <int HDP> struct am_bounds_bwd { static constexpr int min_ctas = HDP <= 16 ? 4 : (HDP <= 32 ? 3 : 2); };
#define LOG2E 1.4426950408889634f

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// shared-window address forms: the per-lane fragment offset is computed once per kernel, tile/step offsets are immediates
__device__ __forceinline__ void ldsm_x4_u(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t_u(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// per-lane smem addresses of the three fragment kinds (LDS = row stride in elements)
template <int LDS> __device__ __forceinline__ const bf16* a_frag_ptr(const bf16* s, int row0, int col0, int lane) {
    return s + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + col0 + (lane >> 4) * 8;
}
// B operand stored [n][k] (k contiguous): two n-tiles (n0..n0+15) x one k-step -> {b0,b1 | b0,b1}
template <int LDS> __device__ __forceinline__ const bf16* b_frag_ptr(const bf16* s, int n0, int k0, int lane) {
    return s + (n0 + (lane & 7) + ((lane >> 4) & 1) * 8) * LDS + k0 + ((lane >> 3) & 1) * 8;
}
// B operand stored [k][n] (n contiguous), read transposed: one k-step (k0..k0+15) x two n-tiles (n0..n0+15)
template <int LDS> __device__ __forceinline__ const bf16* bt_frag_ptr(const bf16* s, int k0, int n0, int lane) {
    return s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + n0 + (lane >> 4) * 8;
}

// global head tile [L rows x HDP cols, 16-byte aligned rows of pitch ld] -> smem [LP][LDS]; rows >= L are zeroed.
// 16-byte cp.async copies: every chunk of every tile of the head is in flight at once (no register staging, no
// per-batch stall on the load latency); the caller waits with cp_async_wait_all() + __syncthreads().
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
template <int HDP, int LDS>
__device__ __forceinline__ void load_head_tile(const bf16* __restrict__ src, int ld, int L, int LP, bf16* __restrict__ dst) {
    // a thread keeps one 16-byte chunk column and walks the rows with constant pointer strides: ~6 instructions per chunk
    // (the index form -- row = i / CPR, column = i % CPR, 64-bit address per chunk -- was 30, 16 % of the forward kernel)
    constexpr int CPR = HDP / 8;                                   // 16-byte chunks per row
    const int c = threadIdx.x % CPR, r0 = threadIdx.x / CPR, rstep = (int)blockDim.x / CPR;   // blockDim is a multiple of 32
    const bf16* s = src + (size_t)r0 * ld + c * 8;
    const size_t sstep = (size_t)rstep * ld;
    uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r0 * LDS + c * 8);
    const uint32_t dstep = (uint32_t)(rstep * LDS * 2);
    int l = r0;
    for (; l < L; l += rstep, s += sstep, d += dstep)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory");
    for (; l < LP; l += rstep, d += dstep)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0u) : "memory");
}
// smem [L][LDS] -> global head tile (128-bit stores, padding columns included: they hold zeros)
template <int HDP, int LDS>
__device__ __forceinline__ void store_head_tile(const bf16* __restrict__ src, bf16* __restrict__ dst, int ld, int L) {
    constexpr int CPR = HDP / 8;
    const int c = threadIdx.x % CPR, r0 = threadIdx.x / CPR, rstep = (int)blockDim.x / CPR;
    const bf16* s = src + r0 * LDS + c * 8;
    bf16* d = dst + (size_t)r0 * ld + c * 8;
    const size_t dstep = (size_t)rstep * ld;
    const int sstep = rstep * LDS;
    for (int l = r0; l < L; l += rstep, s += sstep, d += dstep) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
}
// accumulator tile (16 rows x NTO*8 cols, mma C layout) -> bf16 smem rows [row0, row0+16)
template <int LDS, int NTO>
__device__ __forceinline__ void stage_tile(bf16* __restrict__ dst, int row0, const float (&acc)[NTO][4], float s0, float s1, int g, int t) {
#pragma unroll
    for (int no = 0; no < NTO; ++no) {
        const int col = no * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(dst + (row0 + g) * LDS + col) = __floats2bfloat162_rn(acc[no][0] * s0, acc[no][1] * s0);
        *reinterpret_cast<__nv_bfloat162*>(dst + (row0 + g + 8) * LDS + col) = __floats2bfloat162_rn(acc[no][2] * s1, acc[no][3] * s1);
    }
}

// One key block of the forward pass for one 16-query tile: NP pairs of 8-key n-tiles (16*NP keys) starting at pair pb.
// S = Q K^T -> online softmax update of (m, l, O) -> O += P V.  NP is a template parameter so that the ragged last block
// of a head is a separate straight-line instantiation instead of predicated ldmatrix/mma inside the full-block code.
template <int HDP, int NP>
__device__ __forceinline__ void attn_fwd_block(const bf16* __restrict__ Ks, const bf16* __restrict__ Vs, const uint32_t (&qa)[HDP / 16][4],
                                               float (&oacc)[HDP / 8][4], float& m0, float& m1, float& l0, float& l1, int pb, int L,
                                               float c, int lane, int t) {
    constexpr int LDS = HDP + 8, KS = HDP / 16, NTO = HDP / 8, NT = 2 * NP;
    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t kb[4];
            ldsm_x4(kb, b_frag_ptr<LDS>(Ks, (pb + pp) * 16, ks * 16, lane));
            mma16816(s[2 * pp], qa[ks], kb[0], kb[1]);
            mma16816(s[2 * pp + 1], qa[ks], kb[2], kb[3]);
        }
    }
    if ((pb + NP) * 16 > L) {                                  // only the last key block has columns >= L (warp-uniform)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = pb * 16 + nt * 8 + 2 * t;
            if (col >= L) s[nt][0] = s[nt][2] = -INFINITY;
            if (col + 1 >= L) s[nt][1] = s[nt][3] = -INFINITY;
        }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = fast_exp2((m0 - mn0) * c), al1 = fast_exp2((m1 - mn1) * c);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int i = 0; i < NTO; ++i) { oacc[i][0] *= al0; oacc[i][1] *= al0; oacc[i][2] *= al1; oacc[i][3] *= al1; }
    const float mc0 = mn0 * c, mc1 = mn1 * c;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = fast_exp2(fmaf(s[nt][0], c, -mc0)); s[nt][1] = fast_exp2(fmaf(s[nt][1], c, -mc0));
        s[nt][2] = fast_exp2(fmaf(s[nt][2], c, -mc1)); s[nt][3] = fast_exp2(fmaf(s[nt][3], c, -mc1));
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * pp][0], s[2 * pp][1]);
        pa[1] = pack_bf16(s[2 * pp][2], s[2 * pp][3]);
        pa[2] = pack_bf16(s[2 * pp + 1][0], s[2 * pp + 1][1]);
        pa[3] = pack_bf16(s[2 * pp + 1][2], s[2 * pp + 1][3]);
#pragma unroll
        for (int no = 0; no < NTO; no += 2) {
            uint32_t vb[4];
            ldsm_x4_t(vb, bt_frag_ptr<LDS>(Vs, (pb + pp) * 16, no * 8, lane));
            mma16816(oacc[no], pa, vb[0], vb[1]);
            mma16816(oacc[no + 1], pa, vb[2], vb[3]);
        }
    }
}

template <int HDP>
__global__ void __launch_bounds__(AM_MAX_WARPS * 32, am_bounds<HDP>::min_ctas) attn_fwd_mma_kernel(const bf16* __restrict__ qkv, int ld3, bf16* __restrict__ o,
                                                                     int ldo, float* __restrict__ lse, int L, int d, int H, int halo) {
    constexpr int LDS = HDP + 8, KS = HDP / 16, NTO = HDP / 8;
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int hd = d / H, b = blockIdx.x / H, h = blockIdx.x % H, Lp = L + 2 * halo;
    const int LP = (L + 15) & ~15;
    bf16* Qs = reinterpret_cast<bf16*>(sm_raw);
    bf16* Ks = Qs + LP * LDS;
    bf16* Vs = Ks + LP * LDS;
    const size_t row0 = (size_t)b * Lp + halo;
    const bf16* base = qkv + row0 * ld3 + h * HDP;
    load_head_tile<HDP, LDS>(base, ld3, L, LP, Qs);
    load_head_tile<HDP, LDS>(base + H * HDP, ld3, L, LP, Ks);
    load_head_tile<HDP, LDS>(base + 2 * H * HDP, ld3, L, LP, Vs);
    cp_async_wait_all();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const float sc = rsqrtf((float)hd), c = sc * LOG2E;
    const int npair = LP / 16, nfull = npair & ~3, ntail = npair & 3;
    for (int qt = warp; qt < npair; qt += (int)(blockDim.x >> 5)) {
        uint32_t qa[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) ldsm_x4(qa[ks], a_frag_ptr<LDS>(Qs, qt * 16, ks * 16, lane));
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        float oacc[NTO][4];
#pragma unroll
        for (int i = 0; i < NTO; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
        for (int pb = 0; pb < nfull; pb += 4)                     // 64-key blocks
            attn_fwd_block<HDP, 4>(Ks, Vs, qa, oacc, m0, m1, l0, l1, pb, L, c, lane, t);
        if (ntail == 1) attn_fwd_block<HDP, 1>(Ks, Vs, qa, oacc, m0, m1, l0, l1, nfull, L, c, lane, t);
        else if (ntail == 2) attn_fwd_block<HDP, 2>(Ks, Vs, qa, oacc, m0, m1, l0, l1, nfull, L, c, lane, t);
        else if (ntail == 3) attn_fwd_block<HDP, 3>(Ks, Vs, qa, oacc, m0, m1, l0, l1, nfull, L, c, lane, t);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        const int r0 = qt * 16 + g, r1 = r0 + 8;
        __syncwarp();                                              // all lanes hold their Q fragments of this tile
        stage_tile<LDS, NTO>(Qs, qt * 16, oacc, i0, i1, g, t);    // O overwrites the (now dead) Q rows of this tile
        if (t == 0) {
            if (r0 < L) lse[((size_t)b * H + h) * L + r0] = m0 * sc + __logf(l0);
            if (r1 < L) lse[((size_t)b * H + h) * L + r1] = m1 * sc + __logf(l1);
        }
    }
    __syncthreads();
    store_head_tile<HDP, LDS>(Qs, o + row0 * ldo + h * HDP, ldo, L);
}

__device__ __forceinline__ void am_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void am_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void am_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "AM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra AM_DONE;\n"
        "bra AM_WAIT;\n"
        "AM_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// lane-local column partials (accumulator layout: columns no*8 + 2t, +1; rows g, g+8) -> per-CTA shared sums
template <int NTO>
__device__ __forceinline__ void warp_colsum_flush(float (&c)[NTO][2], float* csum, int g, int t) {
#pragma unroll
    for (int no = 0; no < NTO; ++no) {
        float c0 = c[no][0], c1 = c[no][1];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, off);
            c1 += __shfl_xor_sync(0xffffffffu, c1, off);
        }
        if (g == 0) { atomicAdd(&csum[no * 8 + 2 * t], c0); atomicAdd(&csum[no * 8 + 2 * t + 1], c1); }
    }
}

template <int HDP>
__global__ void __launch_bounds__(AM_MAX_WARPS * 32, am_bounds_bwd<HDP>::min_ctas) attn_bwd_mma_kernel(const bf16* __restrict__ qkv, int ld3, const bf16* __restrict__ o,
                                                                     int ldo, const bf16* __restrict__ dout, int lddo,
                                                                     bf16* __restrict__ dqkv, int lddqkv, const float* __restrict__ lse,
                                                                     int L, int d, int H, int halo, float* __restrict__ dbias) {
    constexpr int LDS = HDP + 8, KS = HDP / 16, NTO = HDP / 8, CPR = HDP / 8;
    __shared__ float csum[3 * HDP];                            // column sums of dQ | dK | dV of this head (in_proj bias gradient)
    __shared__ __align__(8) uint64_t passA_bar;
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int hd = d / H, b = blockIdx.x / H, h = blockIdx.x % H, Lp = L + 2 * halo;
    const int LP = (L + 15) & ~15;
    bf16* Qs = reinterpret_cast<bf16*>(sm_raw);
    bf16* Ks = Qs + LP * LDS;
    bf16* Vs = Ks + LP * LDS;
    bf16* Gs = Vs + LP * LDS;
    float* Ls = reinterpret_cast<float*>(Gs + LP * LDS);      // lse * log2(e)
    float* Ds = Ls + LP;                                      // rowsum(dO * O)
    const size_t row0 = (size_t)b * Lp + halo;
    const bf16* base = qkv + row0 * ld3 + h * HDP;
    load_head_tile<HDP, LDS>(base, ld3, L, LP, Qs);
    load_head_tile<HDP, LDS>(base + H * HDP, ld3, L, LP, Ks);
    load_head_tile<HDP, LDS>(base + 2 * H * HDP, ld3, L, LP, Vs);
    load_head_tile<HDP, LDS>(dout + row0 * lddo + h * HDP, lddo, L, LP, Gs);
    for (int i = threadIdx.x; i < LP; i += blockDim.x) Ls[i] = i < L ? lse[((size_t)b * H + h) * L + i] * LOG2E : 0.f;
    for (int i = threadIdx.x; i < 3 * HDP; i += blockDim.x) csum[i] = 0.f;
    const uint32_t abar = (uint32_t)__cvta_generic_to_shared(&passA_bar);
    if (threadIdx.x == 0) am_mbar_init(abar, blockDim.x >> 5);
    // D_i = rowsum(dO_i * O_i): CPR lanes per row, one 128-bit O load each.  The O chunks of the first NOI rounds are
    // requested BEFORE the wait on the tile copies, so the two global-memory latencies overlap instead of adding up.
    constexpr int NOI = 4;
    const int nchunk = LP * CPR;
    uint4 ov[NOI];
#pragma unroll
    for (int it = 0; it < NOI; ++it) {
        const int i = threadIdx.x + it * blockDim.x, l = i / CPR, cch = i % CPR;
        ov[it] = (i < nchunk && l < L) ? *reinterpret_cast<const uint4*>(o + (row0 + l) * ldo + h * HDP + cch * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    cp_async_wait_all();
    __syncthreads();
    auto d_row = [&](int i, const uint4& ovv) {                   // nchunk and blockDim are multiples of 32: warp-uniform
        const int l = i / CPR, cch = i % CPR;
        const uint4 gv = *reinterpret_cast<const uint4*>(Gs + l * LDS + cch * 8);
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ovv);
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 x = __bfloat1622float2(oh[u]), y = __bfloat1622float2(gh[u]);
            a = fmaf(x.x, y.x, a); a = fmaf(x.y, y.y, a);
        }
#pragma unroll
        for (int off = 1; off < CPR; off <<= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (cch == 0) Ds[l] = a;                                   // rows >= L: O was read as zero
    };
#pragma unroll
    for (int it = 0; it < NOI; ++it) {
        const int i = threadIdx.x + it * blockDim.x;
        if (i < nchunk) d_row(i, ov[it]);
    }
    for (int i = threadIdx.x + NOI * blockDim.x; i < nchunk; i += blockDim.x) {
        const int l = i / CPR, cch = i % CPR;
        const uint4 o4 = l < L ? *reinterpret_cast<const uint4*>(o + (row0 + l) * ldo + h * HDP + cch * 8) : make_uint4(0u, 0u, 0u, 0u);
        d_row(i, o4);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const float sc = rsqrtf((float)hd), c = sc * LOG2E;
    const int ntile = LP / 16;
    // shared-window byte addresses of the fragments of this lane; a tile of 16 rows is TILE_B bytes further on
    constexpr uint32_t TILE_B = 32u * LDS;
    const uint32_t a_off = 2u * (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (lane >> 4) * 8);     // A and transposed-B form
    const uint32_t b_off = 2u * (uint32_t)(((lane & 7) + ((lane >> 4) & 1) * 8) * LDS + ((lane >> 3) & 1) * 8); // B form, [n][k] storage
    const uint32_t Qa = (uint32_t)__cvta_generic_to_shared(Qs) + a_off, Qb = (uint32_t)__cvta_generic_to_shared(Qs) + b_off;
    const uint32_t Ka = (uint32_t)__cvta_generic_to_shared(Ks) + a_off, Kb = (uint32_t)__cvta_generic_to_shared(Ks) + b_off;
    const uint32_t Va = (uint32_t)__cvta_generic_to_shared(Vs) + a_off, Vb = (uint32_t)__cvta_generic_to_shared(Vs) + b_off;
    const uint32_t Ga = (uint32_t)__cvta_generic_to_shared(Gs) + a_off, Gb = (uint32_t)__cvta_generic_to_shared(Gs) + b_off;
    // ---------------- pass A: dQ  (the 1/sqrt(hd) factor of dS is applied once to the finished dQ / dK tiles)
    float cq[NTO][2];
#pragma unroll
    for (int no = 0; no < NTO; ++no) cq[no][0] = cq[no][1] = 0.f;
    for (int qt = warp; qt < ntile; qt += (int)(blockDim.x >> 5)) {
        uint32_t qa[KS][4], ga[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            ldsm_x4_u(qa[ks], Qa + qt * TILE_B + ks * 32);
            ldsm_x4_u(ga[ks], Ga + qt * TILE_B + ks * 32);
        }
        const int r0 = qt * 16 + g, r1 = r0 + 8;
        const float ls0 = Ls[r0], ls1 = Ls[r1], d0 = Ds[r0], d1 = Ds[r1];
        float dq[NTO][4];
#pragma unroll
        for (int i = 0; i < NTO; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
        uint32_t tb = 0;
        for (int kp = 0; kp < ntile; ++kp, tb += TILE_B) {      // 16 keys per iteration
            float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t kb[4], vb[4];
                ldsm_x4_u(kb, Kb + tb + ks * 32);
                ldsm_x4_u(vb, Vb + tb + ks * 32);
                mma16816(s[0], qa[ks], kb[0], kb[1]);
                mma16816(s[1], qa[ks], kb[2], kb[3]);
                mma16816(dp[0], ga[ks], vb[0], vb[1]);
                mma16816(dp[1], ga[ks], vb[2], vb[3]);
            }
            float ds[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float p0 = fast_exp2(s[nt][0] * c - ls0), p1 = fast_exp2(s[nt][1] * c - ls0);
                float p2 = fast_exp2(s[nt][2] * c - ls1), p3 = fast_exp2(s[nt][3] * c - ls1);
                if (kp == ntile - 1) {                             // only the last key tile has columns >= L (warp-uniform)
                    const int col = kp * 16 + nt * 8 + 2 * t;
                    if (col >= L) p0 = p2 = 0.f;
                    if (col + 1 >= L) p1 = p3 = 0.f;
                }
                ds[nt][0] = p0 * (dp[nt][0] - d0); ds[nt][1] = p1 * (dp[nt][1] - d0);
                ds[nt][2] = p2 * (dp[nt][2] - d1); ds[nt][3] = p3 * (dp[nt][3] - d1);
            }
            uint32_t da[4] = {pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]), pack_bf16(ds[1][0], ds[1][1]),
                              pack_bf16(ds[1][2], ds[1][3])};
#pragma unroll
            for (int no = 0; no < NTO; no += 2) {
                uint32_t kb[4];
                ldsm_x4_t_u(kb, Ka + tb + no * 16);
                mma16816(dq[no], da, kb[0], kb[1]);
                mma16816(dq[no + 1], da, kb[2], kb[3]);
            }
        }
#pragma unroll
        for (int no = 0; no < NTO; ++no) { dq[no][0] *= sc; dq[no][1] *= sc; dq[no][2] *= sc; dq[no][3] *= sc; }
        // dQ: bf16x2 stores straight from the accumulator layout (Q and dO tiles are still needed by pass B)
#pragma unroll
        for (int no = 0; no < NTO; ++no) {
            const int col = no * 8 + 2 * t;
            if (r0 < L) *reinterpret_cast<__nv_bfloat162*>(dqkv + (row0 + r0) * lddqkv + h * HDP + col) = __floats2bfloat162_rn(dq[no][0], dq[no][1]);
            if (r1 < L) *reinterpret_cast<__nv_bfloat162*>(dqkv + (row0 + r1) * lddqkv + h * HDP + col) = __floats2bfloat162_rn(dq[no][2], dq[no][3]);
        }
        if (dbias) {                                               // lane-local column sums of the stored (bf16) dQ rows
#pragma unroll
            for (int no = 0; no < NTO; ++no) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(dq[no][0], dq[no][1]), hi = __floats2bfloat162_rn(dq[no][2], dq[no][3]);
                cq[no][0] += (r0 < L ? __low2float(lo) : 0.f) + (r1 < L ? __low2float(hi) : 0.f);
                cq[no][1] += (r0 < L ? __high2float(lo) : 0.f) + (r1 < L ? __high2float(hi) : 0.f);
            }
        }
    }
    if (dbias) warp_colsum_flush<NTO>(cq, csum, g, t);
#if !AM_DIRECT_STORE
    __syncwarp();
    if (lane == 0) am_mbar_arrive(abar);                          // this warp no longer reads K / V rows of other tiles
#endif
    // ---------------- pass B: dK, dV (rows of the accumulators are keys, columns of S^T are queries)
    float ck[NTO][2], cv[NTO][2];
#pragma unroll
    for (int no = 0; no < NTO; ++no) ck[no][0] = ck[no][1] = cv[no][0] = cv[no][1] = 0.f;
    // tiles are dealt in the opposite order to pass A, so a warp with one tile more there has one tile less here
    bool passA_done = false; (void)passA_done;
    for (int kt = ntile - 1 - warp; kt >= 0; kt -= (int)(blockDim.x >> 5)) {
        uint32_t ka[KS][4], va[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            ldsm_x4_u(ka[ks], Ka + kt * TILE_B + ks * 32);
            ldsm_x4_u(va[ks], Va + kt * TILE_B + ks * 32);
        }
        float dk[NTO][4], dv[NTO][4];
#pragma unroll
        for (int i = 0; i < NTO; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
        uint32_t tb = 0;
        for (int qp = 0; qp < ntile; ++qp, tb += TILE_B) {      // 16 queries per iteration
            float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t qb[4], gb[4];
                ldsm_x4_u(qb, Qb + tb + ks * 32);
                ldsm_x4_u(gb, Gb + tb + ks * 32);
                mma16816(st[0], ka[ks], qb[0], qb[1]);
                mma16816(st[1], ka[ks], qb[2], qb[3]);
                mma16816(dpt[0], va[ks], gb[0], gb[1]);
                mma16816(dpt[1], va[ks], gb[2], gb[3]);
            }
            float pt[2][4], dst_[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int q0 = qp * 16 + nt * 8 + 2 * t;
                const float2 lq = *reinterpret_cast<const float2*>(Ls + q0), dd = *reinterpret_cast<const float2*>(Ds + q0);
                pt[nt][0] = fast_exp2(st[nt][0] * c - lq.x); pt[nt][1] = fast_exp2(st[nt][1] * c - lq.y);
                pt[nt][2] = fast_exp2(st[nt][2] * c - lq.x); pt[nt][3] = fast_exp2(st[nt][3] * c - lq.y);
                if (qp == ntile - 1) {                             // only the last query tile has columns >= L (warp-uniform)
                    if (q0 >= L) pt[nt][0] = pt[nt][2] = 0.f;
                    if (q0 + 1 >= L) pt[nt][1] = pt[nt][3] = 0.f;
                }
                dst_[nt][0] = pt[nt][0] * (dpt[nt][0] - dd.x); dst_[nt][1] = pt[nt][1] * (dpt[nt][1] - dd.y);
                dst_[nt][2] = pt[nt][2] * (dpt[nt][2] - dd.x); dst_[nt][3] = pt[nt][3] * (dpt[nt][3] - dd.y);
            }
            uint32_t pa[4] = {pack_bf16(pt[0][0], pt[0][1]), pack_bf16(pt[0][2], pt[0][3]), pack_bf16(pt[1][0], pt[1][1]),
                              pack_bf16(pt[1][2], pt[1][3])};
            uint32_t sa[4] = {pack_bf16(dst_[0][0], dst_[0][1]), pack_bf16(dst_[0][2], dst_[0][3]), pack_bf16(dst_[1][0], dst_[1][1]),
                              pack_bf16(dst_[1][2], dst_[1][3])};
#pragma unroll
            for (int no = 0; no < NTO; no += 2) {
                uint32_t gb[4], qb[4];
                ldsm_x4_t_u(gb, Ga + tb + no * 16);
                ldsm_x4_t_u(qb, Qa + tb + no * 16);
                mma16816(dv[no], pa, gb[0], gb[1]);
                mma16816(dv[no + 1], pa, gb[2], gb[3]);
                mma16816(dk[no], sa, qb[0], qb[1]);
                mma16816(dk[no + 1], sa, qb[2], qb[3]);
            }
        }
#pragma unroll
        for (int no = 0; no < NTO; ++no) { dk[no][0] *= sc; dk[no][1] *= sc; dk[no][2] *= sc; dk[no][3] *= sc; }
#if AM_DIRECT_STORE
        // dK / dV: bf16x2 stores straight from the accumulator layout, like dQ (a quad of lanes writes 16 contiguous bytes of a
        // row).  Nothing is staged over the K / V rows, so the two passes need no barrier between them and the kernel no
        // store loop at its end (ncu: 8 % of the warp samples sat on that barrier).
        {
            const int k0 = kt * 16 + g, k1 = k0 + 8;
            bf16* dkp = dqkv + row0 * lddqkv + H * HDP + h * HDP;
            bf16* dvp = dkp + H * HDP;
#pragma unroll
            for (int no = 0; no < NTO; ++no) {
                const int col = no * 8 + 2 * t;
                if (k0 < L) {
                    *reinterpret_cast<__nv_bfloat162*>(dkp + (size_t)k0 * lddqkv + col) = __floats2bfloat162_rn(dk[no][0], dk[no][1]);
                    *reinterpret_cast<__nv_bfloat162*>(dvp + (size_t)k0 * lddqkv + col) = __floats2bfloat162_rn(dv[no][0], dv[no][1]);
                }
                if (k1 < L) {
                    *reinterpret_cast<__nv_bfloat162*>(dkp + (size_t)k1 * lddqkv + col) = __floats2bfloat162_rn(dk[no][2], dk[no][3]);
                    *reinterpret_cast<__nv_bfloat162*>(dvp + (size_t)k1 * lddqkv + col) = __floats2bfloat162_rn(dv[no][2], dv[no][3]);
                }
            }
        }
#else
        // K/V rows of this tile are otherwise only read as its own A fragments (held in registers by now) and by pass A
        // of the other warps: wait until every warp has left pass A (split barrier: arrived long ago, rarely blocks)
        if (!passA_done) { am_mbar_wait(abar, 0u); passA_done = true; }
        __syncwarp();
        stage_tile<LDS, NTO>(Ks, kt * 16, dk, 1.f, 1.f, g, t);
        stage_tile<LDS, NTO>(Vs, kt * 16, dv, 1.f, 1.f, g, t);
#endif
        if (dbias) {
            const bool v0 = kt * 16 + g < L, v1 = kt * 16 + g + 8 < L;
#pragma unroll
            for (int no = 0; no < NTO; ++no) {
                const __nv_bfloat162 klo = __floats2bfloat162_rn(dk[no][0], dk[no][1]), khi = __floats2bfloat162_rn(dk[no][2], dk[no][3]);
                const __nv_bfloat162 vlo = __floats2bfloat162_rn(dv[no][0], dv[no][1]), vhi = __floats2bfloat162_rn(dv[no][2], dv[no][3]);
                ck[no][0] += (v0 ? __low2float(klo) : 0.f) + (v1 ? __low2float(khi) : 0.f);
                ck[no][1] += (v0 ? __high2float(klo) : 0.f) + (v1 ? __high2float(khi) : 0.f);
                cv[no][0] += (v0 ? __low2float(vlo) : 0.f) + (v1 ? __low2float(vhi) : 0.f);
                cv[no][1] += (v0 ? __high2float(vlo) : 0.f) + (v1 ? __high2float(vhi) : 0.f);
            }
        }
    }
    if (dbias) {
        warp_colsum_flush<NTO>(ck, csum + HDP, g, t);
        warp_colsum_flush<NTO>(cv, csum + 2 * HDP, g, t);
    }
    __syncthreads();
#if !AM_DIRECT_STORE
    store_head_tile<HDP, LDS>(Ks, dqkv + row0 * lddqkv + H * HDP + h * HDP, lddqkv, L);
    store_head_tile<HDP, LDS>(Vs, dqkv + row0 * lddqkv + 2 * H * HDP + h * HDP, lddqkv, L);
#endif
    if (dbias) {                                                   // one atomic per (q|k|v, channel) into the compact bias gradient
        for (int i = threadIdx.x; i < 3 * HDP; i += blockDim.x) {
            const int w = i / HDP, e = i - w * HDP;
            if (e < hd) atomicAdd(dbias + w * d + h * hd + e, csum[i]);
        }
    }
}

// warps per CTA: the 16-row tiles of a head are dealt round-robin, so pick the count that leaves no idle round
static int am_warps(int LP) {
    const int ntile = LP / 16, rounds = (ntile + AM_MAX_WARPS - 1) / AM_MAX_WARPS;
    int w = (ntile + rounds - 1) / rounds;
    if (w > AM_MAX_WARPS) w = AM_MAX_WARPS;
    return w < 1 ? 1 : w;
}
static int pad_hd(int hd) { return hd <= 16 ? 16 : (hd <= 32 ? 32 : 64); }

// eligible: head pitch is exactly the padded MMA width and everything fits in shared memory
extern "C" int csi_attn_mma_ok(int L, int d, int H, int hp) {
    if (H <= 0 || d % H) return 0;
    const int hd = d / H;
    if (hd > 64 || hp != pad_hd(hd)) return 0;
    const int LP = (L + 15) & ~15, LDS = hp + 8;
    return ((size_t)4 * LP * LDS * 2 + 2 * (size_t)LP * 4) <= 200 * 1024;
}

extern "C" int csi_attn_fwd_mma(const void* qkv, int ld3, void* o, int ldo, float* lse, int B, int L, int d, int H, int hp,
                                int halo, void* stream) {
    CSI_CHECK_ARG(qkv && o && lse, "null pointer");
    CSI_CHECK_ARG(csi_attn_mma_ok(L, d, H, hp), "shape not eligible");
    CSI_CHECK_ARG(ld3 % 8 == 0 && ldo % 8 == 0 && ld3 >= 3 * H * hp && ldo >= H * hp, "head-padded leading dimensions expected");
    if (B == 0) return CSI_OK;
    const int LP = (L + 15) & ~15;
    const size_t smem = (size_t)3 * LP * (hp + 8) * 2;
    const int nw = am_warps(LP);
#define GO(HDP)                                                                                                        \
    do {                                                                                                               \
        CSI_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        CSI_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<HDP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        attn_fwd_mma_kernel<HDP><<<B * H, nw * 32, smem, ST(stream)>>>((const bf16*)qkv, ld3, (bf16*)o, ldo, lse, L, d, H, halo); \
    } while (0)
    if (hp == 16) GO(16); else if (hp == 32) GO(32); else GO(64);
#undef GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

extern "C" int csi_attn_bwd_mma(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo, void* dqkv,
                                int lddqkv, const float* lse, int B, int L, int d, int H, int hp, int halo, float* dbias,
                                void* stream) {
    CSI_CHECK_ARG(qkv && o && dout && dqkv && lse, "null pointer");
    CSI_CHECK_ARG(csi_attn_mma_ok(L, d, H, hp), "shape not eligible");
    CSI_CHECK_ARG(ld3 % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddqkv % 8 == 0, "head-padded leading dimensions expected");
    if (B == 0) return CSI_OK;
    const int LP = (L + 15) & ~15;
    const size_t smem = (size_t)4 * LP * (hp + 8) * 2 + 2 * (size_t)LP * 4;
    const int nw = am_warps(LP);
#define GO(HDP)                                                                                                        \
    do {                                                                                                               \
        CSI_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        CSI_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel<HDP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        attn_bwd_mma_kernel<HDP><<<B * H, nw * 32, smem, ST(stream)>>>((const bf16*)qkv, ld3, (const bf16*)o, ldo,       \
                                                                        (const bf16*)dout, lddo, (bf16*)dqkv, lddqkv, lse, L, d, H, halo, dbias); \
    } while (0)
    if (hp == 16) GO(16); else if (hp == 32) GO(32); else GO(64);
#undef GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
