// Bandwidth-bound kernels of the THAT train step (sm_100a): input pooling/augmentation, Gaussian range
// encoding, LayerNorm, BatchNorm+activation, head reductions, loss, Adam, weight re-layout.
// All of them are coalesced along the channel axis, vectorised by 2 (channel counts are even: multiples of 10)
// and reduce with warp shuffles; none synchronises with the host.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void csi_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* csi_last_error(void) { return g_err; }
extern "C" int csi_abi_version(void) { return 5; }      // 5: csi_gemm_nt_banded, csi_gemm_tn_workspace
extern "C" int csi_device_arch(int device) {
    cudaDeviceProp p;
    CSI_CUDA(cudaGetDeviceProperties(&p, device));
    return p.major * 10 + p.minor;
}

#define ST(s) ((cudaStream_t)(s))
static int g_num_sms_ew = 0;
static inline int num_sms() {
    if (g_num_sms_ew == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_num_sms_ew, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            g_num_sms_ew <= 0)
            g_num_sms_ew = 148;
    }
    return g_num_sms_ew;
}
#define SITE_AUG 9001u
#define SITE_AUG_SCALE 9002u

// ------------------------------------------------------------------------------------------------ 8-wide access
// 8 consecutive channels per thread: one 16-byte load for bf16, two for fp32 (pointers are 16-byte aligned because
// every leading dimension / column offset is a multiple of 8 elements).
template <typename T> __device__ __forceinline__ void load8f(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8f<float>(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8f<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store8f(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8f<float>(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8f<bf16>(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// raw (unconverted) 8-channel chunk: lets a loop issue the next row's loads before converting the current one
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 v;
    __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void get(float (&f)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 x = __bfloat1622float2(h[i]); f[2 * i] = x.x; f[2 * i + 1] = x.y; }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = reinterpret_cast<const float4*>(p)[0]; b = reinterpret_cast<const float4*>(p)[1]; }
    __device__ __forceinline__ void get(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};

// valid token r = (sample b, position l) -> physical row b*Lp + halo + l, advanced without divisions
struct RowWalk {
    int b, l, L, Lp, halo;
    __device__ __forceinline__ void init(int r, int L_, int halo_) { L = L_; halo = halo_; Lp = L_ + 2 * halo_; b = r / L_; l = r - b * L_; }
    __device__ __forceinline__ size_t row() const { return (size_t)b * Lp + halo + l; }
    __device__ __forceinline__ void advance(int step) { l += step; while (l >= L) { l -= L; ++b; } }
};

// ------------------------------------------------------------------------------------------------ pool_dual
// One CTA = (sample, 16 pooled tokens).  A work item is (token, 4 consecutive features): 20 x 2 float2 loads, one
// Philox call per (time step, 4 features) when augmenting (2 Box-Muller pairs from 16-bit uniforms + 4 keep bits).
#define POOL_K 20
// 8 tokens per CTA: 19 x B CTAs of 288 threads at L=150, F=270 (every thread owns 2 of the 544 work items; ~8 waves).  With 16
// tokens the 2560 CTAs were 2.2 waves of 8 resident CTAs and each thread looped 4.25 -> 5 times: 61 % of the machine.
// 8 tokens = 32 bytes per row of the right-stream layout, one whole sector per store.
#define POOL_TL 8

template <bool AUG>
__global__ void __launch_bounds__(512) pool_dual_kernel(
    const float* __restrict__ x, const long long* __restrict__ offs, const int* __restrict__ lens, int T, int F,
    const float* __restrict__ pe, int ld_pe, float* __restrict__ left, int ld_left, float* __restrict__ right,
    int ld_right, int halo, const unsigned long long* __restrict__ rng) {
    extern __shared__ float tile[];                    // [POOL_TL][F + 1]
    const int b = blockIdx.y, l0 = blockIdx.x * POOL_TL, L = T / POOL_K;
    const int ntl = min(POOL_TL, L - l0);
    const int Lp_l = L + 2 * halo, Lp_r = F + 2 * halo, FS = F + 1, G = (F + 3) >> 2;
    const float* xb;
    int pad = 0;
    if (offs) { xb = x + offs[b]; pad = T - lens[b]; } else { xb = x + (size_t)b * T * F; }
    RngKey rk;
    float scale = 1.f;
    if (AUG) {
        rk = rng_load(rng);
        const uint4 g = rng_group(rk, SITE_AUG_SCALE, (unsigned long long)b);
        scale = (float)g.x * (0.2f / 4294967296.0f) + 0.9f;           // U[0.9, 1.1)
    }
    const uint32_t keep_thr = 2621;                                   // Bernoulli(0.96): drop iff u16 < 0.04 * 65536
    for (int w = threadIdx.x; w < ntl * G; w += blockDim.x) {
        const int tl = w / G, gq = w % G, l = l0 + tl, f0 = gq * 4;
        const bool hi = (f0 + 2) < F;                                 // F is even: a group holds 4 or 2 valid features
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (!AUG) {
#pragma unroll 4
            for (int i = 0; i < POOL_K; ++i) {
                const int tau = l * POOL_K + i;
                if (tau >= pad) {
                    const float* src = xb + (size_t)(tau - pad) * F + f0;
                    const float2 v0 = *reinterpret_cast<const float2*>(src);
                    acc[0] += v0.x; acc[1] += v0.y;
                    if (hi) { const float2 v1 = *reinterpret_cast<const float2*>(src + 2); acc[2] += v1.x; acc[3] += v1.y; }
                }
            }
        } else {
            // train.py:65-73 fused with the pooling that consumes it:  mean_i[(x_i + 0.1 n_i) * s * m_i]
            //   = (s/20) * sum_i m_i x_i  +  (0.1 s/20) * sum_i m_i n_i,   and  sum_i m_i n_i | m  ~  N(0, #kept)
            // so one normal per pooled output (scaled by sqrt(#kept)) is exactly equal in distribution to 20 of them.
            // Keep masks: one Philox call = 8 x 16-bit fields = 2 time steps x 4 features.
            const unsigned long long cbase = (((unsigned long long)b * L + l) * G + gq) * 16;
            float kept[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
            for (int c = 0; c < POOL_K / 2; ++c) {
                const uint4 g = rng_group(rk, SITE_AUG, cbase + c);
#pragma unroll
                for (int hstep = 0; hstep < 2; ++hstep) {
                    const int tau = l * POOL_K + 2 * c + hstep;
                    const uint32_t w0 = hstep ? g.z : g.x, w1 = hstep ? g.w : g.y;
                    const float m0 = (w0 & 0xFFFFu) >= keep_thr ? 1.f : 0.f, m1 = (w0 >> 16) >= keep_thr ? 1.f : 0.f;
                    const float m2 = (w1 & 0xFFFFu) >= keep_thr ? 1.f : 0.f, m3 = (w1 >> 16) >= keep_thr ? 1.f : 0.f;
                    kept[0] += m0; kept[1] += m1; kept[2] += m2; kept[3] += m3;
                    if (tau >= pad) {
                        const float* src = xb + (size_t)(tau - pad) * F + f0;
                        const float2 v0 = *reinterpret_cast<const float2*>(src);
                        acc[0] += v0.x * m0; acc[1] += v0.y * m1;
                        if (hi) { const float2 v1 = *reinterpret_cast<const float2*>(src + 2); acc[2] += v1.x * m2; acc[3] += v1.y * m3; }
                    }
                }
            }
            const uint4 g = rng_group(rk, SITE_AUG, cbase + 15);
            const float ua = ((float)(g.x >> 8) + 1.0f) * (1.0f / 16777216.0f), ub = (float)(g.y >> 8) * (1.0f / 16777216.0f);
            const float uc = ((float)(g.z >> 8) + 1.0f) * (1.0f / 16777216.0f), ud = (float)(g.w >> 8) * (1.0f / 16777216.0f);
            const float ra = sqrtf(-2.0f * __logf(ua)) * 0.1f, rc = sqrtf(-2.0f * __logf(uc)) * 0.1f;
            float s0, c0, s1, c1;
            __sincosf(6.283185307179586f * ub, &s0, &c0);
            __sincosf(6.283185307179586f * ud, &s1, &c1);
            acc[0] = (acc[0] + ra * c0 * sqrtf(kept[0])) * scale;
            acc[1] = (acc[1] + ra * s0 * sqrtf(kept[1])) * scale;
            acc[2] = (acc[2] + rc * c1 * sqrtf(kept[2])) * scale;
            acc[3] = (acc[3] + rc * s1 * sqrtf(kept[3])) * scale;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] *= (1.0f / POOL_K);
        float* lrow = left + ((size_t)b * Lp_l + halo + l) * ld_left + f0;
        tile[tl * FS + f0] = acc[0];
        tile[tl * FS + f0 + 1] = acc[1];
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (pe) { p0 = pe[l * ld_pe + f0]; p1 = pe[l * ld_pe + f0 + 1]; }
        *reinterpret_cast<float2*>(lrow) = make_float2(acc[0] + p0, acc[1] + p1);
        if (hi) {
            tile[tl * FS + f0 + 2] = acc[2];
            tile[tl * FS + f0 + 3] = acc[3];
            if (pe) { p2 = pe[l * ld_pe + f0 + 2]; p3 = pe[l * ld_pe + f0 + 3]; }
            *reinterpret_cast<float2*>(lrow + 2) = make_float2(acc[2] + p2, acc[3] + p3);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < F * POOL_TL; idx += blockDim.x) {
        const int f = idx / POOL_TL, tl = idx % POOL_TL;
        if (tl < ntl) right[((size_t)b * Lp_r + halo + f) * ld_right + l0 + tl] = tile[tl * FS + f];
    }
}

// ------------------------------------------------------------------------------------------------ gaussian PE
__global__ void gauss_pe_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ mu,
                                    const float* __restrict__ sigma, const float* __restrict__ emb, int K, int F,
                                    float* __restrict__ w, float* __restrict__ pe, int ld_pe) {
    __shared__ float sw[64];
    const int l = blockIdx.x;
    if (threadIdx.x == 0) {
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
            float df = pos[l * K + k] - mu[k];
            float lp = -(df * df) / sigma[k] / sigma[k] / 2.f - logf(sigma[k]);
            sw[k] = lp; mx = fmaxf(mx, lp);
        }
        float s = 0.f;
        for (int k = 0; k < K; ++k) { sw[k] = expf(sw[k] - mx); s += sw[k]; }
        for (int k = 0; k < K; ++k) { sw[k] /= s; w[l * K + k] = sw[k]; }
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < K; ++k) a += sw[k] * emb[k * F + f];
        pe[l * ld_pe + f] = a;
    }
}

extern "C" int csi_gauss_pe_fwd(const float* pos, const float* mu, const float* sigma, const float* emb, int L,
                                int K, int F, float* w, float* pe, int ld_pe, void* stream) {
    CSI_CHECK_ARG(pos && mu && sigma && emb && w && pe, "null pointer");
    CSI_CHECK_ARG(K <= 64, "at most 64 gaussians");
    gauss_pe_fwd_kernel<<<L, 128, 0, ST(stream)>>>(pos, mu, sigma, emb, K, F, w, pe, ld_pe);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// dpe[l, f] = sum_b dleft[row(b,l), f]; batch split over blockIdx.z with atomics into a zeroed scratch
__global__ void batch_sum_kernel(const float* __restrict__ dleft, int ld, int B, int L, int F, int halo,
                                 float* __restrict__ dpe, int ld_ws, int bchunk) {
    const int l = blockIdx.x, f = blockIdx.y * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int Lp = L + 2 * halo;
    const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
    float a = 0.f;
    for (int b = b0; b < b1; ++b) a += dleft[((size_t)b * Lp + halo + l) * ld + f];
    atomicAdd(dpe + l * ld_ws + f, a);
}

__global__ void gauss_pe_bwd_kernel(const float* __restrict__ dpe, int ld_ws, const float* __restrict__ w,
                                    const float* __restrict__ pos, const float* __restrict__ mu,
                                    const float* __restrict__ sigma, const float* __restrict__ emb, int K, int F,
                                    float* __restrict__ demb, float* __restrict__ dmu, float* __restrict__ dsigma) {
    __shared__ float dw[64];
    __shared__ float red[4];
    const int l = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;   // 128 threads
    for (int k = 0; k < K; ++k) {
        float a = 0.f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) a += dpe[l * ld_ws + f] * emb[k * F + f];
        a = warp_sum(a);
        if (lane == 0) red[wid] = a;
        __syncthreads();
        if (threadIdx.x == 0) dw[k] = red[0] + red[1] + red[2] + red[3];
        __syncthreads();
    }
    for (int k = 0; k < K; ++k) {
        const float wk = w[l * K + k];
        for (int f = threadIdx.x; f < F; f += blockDim.x) atomicAdd(demb + k * F + f, wk * dpe[l * ld_ws + f]);
    }
    if (threadIdx.x == 0) {
        float dot = 0.f;
        for (int k = 0; k < K; ++k) dot += w[l * K + k] * dw[k];
        for (int k = 0; k < K; ++k) {
            float dl = w[l * K + k] * (dw[k] - dot);
            float df = pos[l * K + k] - mu[k], sg = sigma[k];
            atomicAdd(dmu + k, dl * df / (sg * sg));
            atomicAdd(dsigma + k, dl * (df * df / (sg * sg * sg) - 1.f / sg));
        }
    }
}

extern "C" int csi_gauss_pe_bwd(const float* dleft, int ld_dleft, int B, int halo, const float* w, const float* pos,
                                const float* mu, const float* sigma, const float* emb, int L, int K, int F,
                                float* dpe_ws, int ld_ws, float* demb, float* dmu, float* dsigma, void* stream) {
    CSI_CHECK_ARG(dleft && w && pos && mu && sigma && emb && dpe_ws && demb && dmu && dsigma, "null pointer");
    CSI_CHECK_ARG(K <= 64, "at most 64 gaussians");
    CSI_CUDA(cudaMemsetAsync(dpe_ws, 0, (size_t)L * ld_ws * sizeof(float), ST(stream)));
    const int bchunk = 16;
    dim3 grid(L, cdiv(F, 128), cdiv(B, bchunk));
    batch_sum_kernel<<<grid, 128, 0, ST(stream)>>>(dleft, ld_dleft, B, L, F, halo, dpe_ws, ld_ws, bchunk);
    CSI_LAUNCH_CHECK();
    gauss_pe_bwd_kernel<<<L, 128, 0, ST(stream)>>>(dpe_ws, ld_ws, w, pos, mu, sigma, emb, K, F, demb, dmu, dsigma);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ layernorm
#define LNB2_WARPS 8
__device__ __forceinline__ uint32_t ew_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ew_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ew_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ew_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ew_mbar_wait(uint32_t bar, uint32_t parity) {
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
// Two variants were measured at B=256, F=270 against the 295 us of the kernel above and removed: streaming the input
// through a cp.async.bulk shared-memory ring four tokens ahead (359 us: 10 instead of 35 resident warps per SM), and placing
// the dropped cells with geometric gaps instead of one 16-bit Philox field per cell (316 us: a third of the instructions,
// but a divergent loop).  Neither load latency nor instruction count was the limit; the grid shape was (see csi_pool_dual).
extern "C" int csi_pool_dual(const float* x, const long long* offs, const int* lens, int B, int T, int F,
                             const float* pe, int ld_pe, float* left, int ld_left, float* right, int ld_right,
                             int halo, int augment, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(x && left && right, "null pointer");
    CSI_CHECK_ARG(T % POOL_K == 0 && F % 2 == 0 && F <= 2800, "T must be a multiple of 20, F even and <= 2800");
    CSI_CHECK_ARG((offs == nullptr) == (lens == nullptr), "offs and lens go together");
    CSI_CHECK_ARG(!augment || rng, "augmentation needs rng");
    if (B == 0) return CSI_OK;
    const int L = T / POOL_K;
    dim3 grid(cdiv(L, POOL_TL), B);
    const size_t smem = (size_t)POOL_TL * (F + 1) * sizeof(float);
    const int items = POOL_TL * ((F + 3) / 4), rounds = cdiv(items, 384);       // whole rounds of work items per thread
    int threads = ((cdiv(items, rounds) + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    if (augment) {
        if (smem > 48 * 1024) CSI_CUDA(cudaFuncSetAttribute(pool_dual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pool_dual_kernel<true><<<grid, threads, smem, ST(stream)>>>(x, offs, lens, T, F, pe, ld_pe, left, ld_left, right, ld_right, halo, rng);
    } else {
        if (smem > 48 * 1024) CSI_CUDA(cudaFuncSetAttribute(pool_dual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pool_dual_kernel<false><<<grid, threads, smem, ST(stream)>>>(x, offs, lens, T, F, pe, ld_pe, left, ld_left, right, ld_right, halo, rng);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

template <typename T> __device__ __forceinline__ float4 lds4f(const uint8_t* p);
template <> __device__ __forceinline__ float4 lds4f<float>(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 lds4f<bf16>(const uint8_t* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void stg4f(T* p, float4 v);
template <> __device__ __forceinline__ void stg4f<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void stg4f<bf16>(bf16* p, float4 v) {
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

// one warp per token row; NP = float2 pairs per lane (d <= 64*NP)
template <typename TY, int NP>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, int ldx,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     TY* __restrict__ y, int ldy, float* __restrict__ mean,
                                                     float* __restrict__ rstd, int B, int L, int d, int halo, float eps) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int Lp = L + 2 * halo, np = d >> 1;
    const float inv_d = 1.0f / d;
    for (int r = gw; r < B * L; r += nw) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        float2 v[NP];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            v[j] = p < np ? *reinterpret_cast<const float2*>(x + row * ldx + 2 * p) : make_float2(0.f, 0.f);
            s += v[j].x + v[j].y;
        }
        const float mu = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) { float a = v[j].x - mu, c = v[j].y - mu; q += a * a + c * c; }
        }
        const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) {
                float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p);
                float2 bb = *reinterpret_cast<const float2*>(beta + 2 * p);
                st2<TY>(y + row * ldy + 2 * p, make_float2((v[j].x - mu) * rs * g.x + bb.x, (v[j].y - mu) * rs * g.y + bb.y));
            }
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    }
}

template <typename TY>
static int ln_fwd_launch(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                         float* mean, float* rstd, int B, int L, int d, int halo, float eps, cudaStream_t s) {
    const int rows = B * L;
    int blocks = cdiv(rows, 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
#define LN_CASE(NP) ln_fwd_kernel<TY, NP><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, (TY*)y, ldy, mean, rstd, B, L, d, halo, eps)
    if (d <= 192) LN_CASE(3); else if (d <= 320) LN_CASE(5); else if (d <= 576) LN_CASE(9); else LN_CASE(16);
#undef LN_CASE
    return 0;
}

extern "C" int csi_layernorm_fwd(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                                 int y_dtype, float* mean, float* rstd, int B, int L, int d, int halo, float eps,
                                 void* stream) {
    CSI_CHECK_ARG(x && gamma && beta && y && mean && rstd, "null pointer");
    CSI_CHECK_ARG(d % 2 == 0 && d <= 1024 && ldx % 2 == 0 && ldy % 2 == 0, "d must be even and <= 1024");
    if (B * L == 0) return CSI_OK;
    if (y_dtype == CSI_BF16) ln_fwd_launch<bf16>(x, ldx, gamma, beta, y, ldy, mean, rstd, B, L, d, halo, eps, ST(stream));
    else ln_fwd_launch<float>(x, ldx, gamma, beta, y, ldy, mean, rstd, B, L, d, halo, eps, ST(stream));
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ layernorm backward
// warp per row, float2 per lane, the next row's operands are fetched before the current row is reduced.
template <typename TDY, typename TM, int NP>
__global__ void __launch_bounds__(256) ln_bwd_kernel(
    const TDY* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres, int lddres,
    float* __restrict__ dx, int lddx, TM* __restrict__ dxm, int lddxm, float drop_p, unsigned drop_site,
    const unsigned long long* __restrict__ rng, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int L,
    int d, int halo) {
    __shared__ float2 sg[8][NP * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int Lp = L + 2 * halo, np = d >> 1, ld8 = ((d + 15) & ~15) >> 3;
    const float inv_d = 1.0f / d;
    const bool drop = (dxm != nullptr) && drop_p > 0.f;
    DropCtx dc;
    if (drop) dc = drop_ctx(rng, drop_p);
    float2 ag[NP], ab[NP], gm[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        ag[j] = ab[j] = make_float2(0.f, 0.f);
        const int p = lane + 32 * j;
        gm[j] = p < np ? *reinterpret_cast<const float2*>(gamma + 2 * p) : make_float2(0.f, 0.f);
    }
    const int total = B * L;
    float2 dvn[NP], xvn[NP], rvn[NP];
    float mun = 0.f, rsn = 0.f;
    auto fetch = [&](int r) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        mun = mean[row]; rsn = rstd[row];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) {
                dvn[j] = ld2<TDY>(dy + row * lddy + 2 * p);
                xvn[j] = *reinterpret_cast<const float2*>(x + row * ldx + 2 * p);
                rvn[j] = dres ? *reinterpret_cast<const float2*>(dres + row * lddres + 2 * p) : make_float2(0.f, 0.f);
            } else { dvn[j] = xvn[j] = rvn[j] = make_float2(0.f, 0.f); }
        }
    };
    constexpr bool PREFETCH = NP <= 5;                 // wider rows would spill with two rows in flight
    if (PREFETCH && gw < total) fetch(gw);
    for (int r = gw; r < total; r += nw) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        if (!PREFETCH) fetch(r);
        float2 dv[NP], xv[NP], rv[NP];
        const float mu = mun, rs = rsn;
#pragma unroll
        for (int j = 0; j < NP; ++j) { dv[j] = dvn[j]; xv[j] = xvn[j]; rv[j] = rvn[j]; }
        if (PREFETCH && r + nw < total) fetch(r + nw);
        float2 g[NP], xh[NP];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            xh[j] = make_float2((xv[j].x - mu) * rs, (xv[j].y - mu) * rs);
            g[j] = make_float2(dv[j].x * gm[j].x, dv[j].y * gm[j].y);
            s1 += g[j].x + g[j].y;
            s2 += g[j].x * xh[j].x + g[j].y * xh[j].y;
            if (lane + 32 * j < np) {
                ag[j].x += dv[j].x * xh[j].x; ag[j].y += dv[j].y * xh[j].y;
                ab[j].x += dv[j].x; ab[j].y += dv[j].y;
            }
        }
        s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) {
                float2 o = make_float2(rs * (g[j].x - s1 - xh[j].x * s2) + rv[j].x, rs * (g[j].y - s1 - xh[j].y * s2) + rv[j].y);
                *reinterpret_cast<float2*>(dx + row * lddx + 2 * p) = o;
                if (dxm) {
                    if (drop) {
                        const uint4 gq = rng_group(dc.k, drop_site, (unsigned long long)row * ld8 + (p >> 2));
                        const int j0 = (2 * p) & 7;
                        o.x *= field16(gq, j0) >= dc.thr ? dc.inv_keep : 0.f;
                        o.y *= field16(gq, j0 + 1) >= dc.thr ? dc.inv_keep : 0.f;
                    }
                    st2<TM>(dxm + row * lddxm + 2 * p, o);
                }
            }
        }
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int j = 0; j < NP; ++j) sg[wid][lane + 32 * j] = pass == 0 ? ag[j] : ab[j];
        __syncthreads();
        for (int p = threadIdx.x; p < np; p += blockDim.x) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 8; ++w) { a.x += sg[w][p].x; a.y += sg[w][p].y; }
            float* dst = pass == 0 ? dgamma : dbeta;
            atomicAdd(dst + 2 * p, a.x); atomicAdd(dst + 2 * p + 1, a.y);
        }
        __syncthreads();
    }
}

// ---- LayerNorm backward, staged version: every warp streams its rows through shared memory with 1-D bulk copies
// (cp.async.bulk + mbarrier, `stages` rows in flight per warp), so the memory-level parallelism does not cost registers
// and wide rows (d = 540) prefetch as deeply as narrow ones.  A lane owns the float4 column groups lane, lane+32, ...
// Grid = 2 CTAs per SM, each warp walks one contiguous slab of rows; dgamma/dbeta are reduced once per CTA.
template <typename TDY, typename TM, int NQ>
__global__ void __launch_bounds__(LNB2_WARPS * 32, 2) ln_bwd2_kernel(
    const TDY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
    float* __restrict__ dx, TM* __restrict__ dxm, float drop_p, unsigned drop_site,
    const unsigned long long* __restrict__ rng, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int L,
    int d, int halo, int ld, int rows_per_warp, int stages, int R) {
    // every buffer has the same row pitch ld (= cols): R consecutive PHYSICAL rows (halo rows included, they are
    // skipped when computing) are one contiguous block, fetched with one bulk copy per operand
    extern __shared__ __align__(16) uint8_t lsm[];
    __shared__ __align__(8) uint64_t bars[LNB2_WARPS * 4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cols = ld, nq = cols >> 2, ld8 = ((d + 15) & ~15) >> 3, Lp = L + 2 * halo;
    const int prows = B * Lp;                                     // physical rows
    const int gw = blockIdx.x * LNB2_WARPS + wid;
    const int p0 = min(prows, gw * rows_per_warp), p1 = min(prows, p0 + rows_per_warp);
    const int n = (p1 - p0 + R - 1) / R;                          // chunks of this warp
    const uint32_t x_bytes = (uint32_t)cols * 4u * R, dy_bytes = (uint32_t)cols * (uint32_t)sizeof(TDY) * R;
    const uint32_t slot_bytes = 2u * x_bytes + dy_bytes;          // [x][dres][dy], R rows each
    const uint32_t tx_bytes = x_bytes + dy_bytes + (dres ? x_bytes : 0u);
    uint8_t* wbase = lsm + (size_t)wid * stages * slot_bytes;
    const uint32_t wbase_u = ew_smem_u32(wbase), bar_u = ew_smem_u32(&bars[wid * 4]);
    if (lane == 0)
        for (int s2 = 0; s2 < stages; ++s2) ew_mbar_init(bar_u + 8u * s2, 1);
    const float inv_d = 1.0f / d;
    const bool drop = (dxm != nullptr) && drop_p > 0.f;
    DropCtx dc;
    if (drop) dc = drop_ctx(rng, drop_p);
    // gamma (zero beyond d) lives in shared memory behind the staging slots; dgamma/dbeta partials stay in registers
    float* sgam = reinterpret_cast<float*>(lsm + (size_t)LNB2_WARPS * stages * slot_bytes);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) sgam[c] = c < d ? gamma[c] : 0.f;
    float4 ag[NQ], ab[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int si = 0, sc = 0, ki = 0;
    uint32_t pc = 0;
    auto issue = [&]() {                                          // chunk ki -> slot si (rows past the buffer end are guard rows)
        if (lane == 0) {
            const size_t row = (size_t)p0 + (size_t)ki * R;
            const uint32_t dst = wbase_u + (uint32_t)si * slot_bytes, bar = bar_u + 8u * si;
            ew_mbar_expect_tx(bar, tx_bytes);
            ew_bulk_g2s(dst, x + row * ld, x_bytes, bar);
            if (dres) ew_bulk_g2s(dst + x_bytes, dres + row * ld, x_bytes, bar);
            ew_bulk_g2s(dst + 2u * x_bytes, dy + row * ld, dy_bytes, bar);
        }
        ++ki;
        if (++si == stages) si = 0;
    };
    const int npro = min(stages, n);
    for (int k = 0; k < npro; ++k) issue();
    int lpos = p0 % Lp;                                           // position of the current physical row inside its sample
    for (int k = 0; k < n; ++k) {
        ew_mbar_wait(bar_u + 8u * sc, pc);
        const uint8_t* slot = wbase + (size_t)sc * slot_bytes;
        for (int rr = 0; rr < R; ++rr) {
            const int prow = p0 + k * R + rr;
            const bool valid = prow < p1 && lpos >= halo && lpos < halo + L;
            if (++lpos == Lp) lpos = 0;
            if (!valid) continue;                                 // warp-uniform
            const size_t row = (size_t)prow;
            const float mu = mean[row], rs = rstd[row];
            const uint8_t* xs = slot + (size_t)rr * cols * 4;
            const uint8_t* rsd = slot + x_bytes + (size_t)rr * cols * 4;
            const uint8_t* ds = slot + 2u * x_bytes + (size_t)rr * cols * sizeof(TDY);
            float4 xh[NQ], dv[NQ];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int qd = lane + 32 * j;
                if (qd < nq) {
                    const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)qd * 16);
                    const float4 gm = *reinterpret_cast<const float4*>(sgam + 4 * qd);
                    dv[j] = lds4f<TDY>(ds + (size_t)qd * 4 * sizeof(TDY));
                    xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                    const float4 g = make_float4(dv[j].x * gm.x, dv[j].y * gm.y, dv[j].z * gm.z, dv[j].w * gm.w);
                    s1 += (g.x + g.y) + (g.z + g.w);
                    s2 += (g.x * xh[j].x + g.y * xh[j].y) + (g.z * xh[j].z + g.w * xh[j].w);
                } else {
                    xh[j] = dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
            // Dropout decisions of the masked copy: an 8-channel Philox group serves the float4 column groups of TWO
            // neighbouring lanes, and a lane meets two groups in two consecutive j.  The even lane of a pair generates the
            // group of iteration j, the odd lane the one of j+1, and they swap the halves they need (two shuffles): one
            // Philox call per lane and pair of iterations instead of two, same decisions bit for bit.
            uint2 kw[NQ];
            if (drop) {
                const bool odd = (lane & 1) != 0;
                const unsigned long long g0 = (unsigned long long)row * ld8 + (lane >> 1);
#pragma unroll
                for (int j = 0; j < NQ; j += 2) {
                    if (j + 1 < NQ) {
                        const uint4 g = rng_group(dc.k, drop_site, g0 + 16 * (odd ? j + 1 : j));
                        const uint32_t t0 = __shfl_xor_sync(0xffffffffu, odd ? g.x : g.z, 1);
                        const uint32_t t1 = __shfl_xor_sync(0xffffffffu, odd ? g.y : g.w, 1);
                        const uint2 own = odd ? make_uint2(g.z, g.w) : make_uint2(g.x, g.y), got = make_uint2(t0, t1);
                        kw[j] = odd ? got : own;
                        kw[j + 1] = odd ? own : got;
                    } else {
                        const uint4 g = rng_group(dc.k, drop_site, g0 + 16 * j);
                        kw[j] = odd ? make_uint2(g.z, g.w) : make_uint2(g.x, g.y);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int qd = lane + 32 * j, c = 4 * qd;
                if (qd < nq) {
                    const float4 gm = *reinterpret_cast<const float4*>(sgam + c);
                    const float4 rv = dres ? *reinterpret_cast<const float4*>(rsd + (size_t)qd * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                    // gamma is zero beyond d, so g is; masking dv keeps stray pad values out of dgamma / dbeta
                    const bool k0 = c < d, k1 = c + 1 < d, k2 = c + 2 < d, k3 = c + 3 < d;
                    ag[j].x += k0 ? dv[j].x * xh[j].x : 0.f; ag[j].y += k1 ? dv[j].y * xh[j].y : 0.f;
                    ag[j].z += k2 ? dv[j].z * xh[j].z : 0.f; ag[j].w += k3 ? dv[j].w * xh[j].w : 0.f;
                    ab[j].x += k0 ? dv[j].x : 0.f; ab[j].y += k1 ? dv[j].y : 0.f;
                    ab[j].z += k2 ? dv[j].z : 0.f; ab[j].w += k3 ? dv[j].w : 0.f;
                    float4 o;
                    o.x = k0 ? rs * (dv[j].x * gm.x - s1 - xh[j].x * s2) + rv.x : 0.f;
                    o.y = k1 ? rs * (dv[j].y * gm.y - s1 - xh[j].y * s2) + rv.y : 0.f;
                    o.z = k2 ? rs * (dv[j].z * gm.z - s1 - xh[j].z * s2) + rv.z : 0.f;
                    o.w = k3 ? rs * (dv[j].w * gm.w - s1 - xh[j].w * s2) + rv.w : 0.f;
                    stg4f<float>(dx + row * ld + c, o);
                    if (dxm) {
                        if (drop) {
                            o.x *= (kw[j].x & 0xFFFFu) >= dc.thr ? dc.inv_keep : 0.f;
                            o.y *= (kw[j].x >> 16) >= dc.thr ? dc.inv_keep : 0.f;
                            o.z *= (kw[j].y & 0xFFFFu) >= dc.thr ? dc.inv_keep : 0.f;
                            o.w *= (kw[j].y >> 16) >= dc.thr ? dc.inv_keep : 0.f;
                        }
                        stg4f<TM>(dxm + row * ld + c, o);
                    }
                }
            }
        }
        __syncwarp();                                              // every lane has read its part of the slot: refill it
        if (k + stages < n) issue();
        if (++sc == stages) { sc = 0; pc ^= 1u; }
    }
    // per-CTA reduction of dgamma / dbeta through the (now idle) staging memory, then one atomic per column
    __syncthreads();
    float4* red = reinterpret_cast<float4*>(lsm);                  // [LNB2_WARPS][nq]
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const int qd = lane + 32 * j;
            if (qd < nq) red[wid * nq + qd] = pass == 0 ? ag[j] : ab[j];
        }
        __syncthreads();
        float* dst = pass == 0 ? dgamma : dbeta;
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < LNB2_WARPS; ++w) a += reinterpret_cast<const float*>(red + w * nq)[c];
            atomicAdd(dst + c, a);
        }
        __syncthreads();
    }
}

template <typename TDY, typename TM>
static void ln_bwd_launch(const void* dy, int lddy, const float* x, int ldx, const float* gamma, const float* mean,
                          const float* rstd, const float* dres, int lddres, float* dx, int lddx, void* dxm, int lddxm,
                          float drop_p, unsigned site, const unsigned long long* rng, float* dgamma, float* dbeta,
                          int B, int L, int d, int halo, cudaStream_t s) {
    // staged kernel: every operand shares one 16-byte aligned row pitch (true for the engine's token buffers)
    const int ld = ldx;
    const bool same = lddy == ld && lddx == ld && (!dres || lddres == ld) && (!dxm || lddxm == ld) && ld % 8 == 0 && ld >= d &&
                      ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
                      (!dres || (uintptr_t)dres % 16 == 0) && (!dxm || (uintptr_t)dxm % 8 == 0);
    if (same && ld <= 640 && B * L >= 64 && halo <= 16) {
        const int prows = B * (L + 2 * halo);
        const size_t rowb = (size_t)2 * ld * 4 + (size_t)ld * sizeof(TDY);
        // R rows per bulk copy, `stages` copies in flight per warp, <= ~100 KB per CTA (two CTAs per SM)
        int R = 2, stages = 2;                         // measured: 2 x 2 ~ 1 x 3 > 4 x 1 (d = 270); what matters is ~90 KB in flight
        if (const char* e = getenv("CSI_LN_R")) R = atoi(e);
        if (const char* e = getenv("CSI_LN_STAGES")) stages = atoi(e);
        if (R < 1 || R > 4 || stages < 1 || stages > 4) { R = 2; stages = 2; }
        while (R > 1 && (size_t)LNB2_WARPS * stages * R * rowb > 100 * 1024) R >>= 1;
        while (stages > 1 && (size_t)LNB2_WARPS * stages * R * rowb > 100 * 1024) --stages;
        const size_t smem = (size_t)LNB2_WARPS * stages * R * rowb + (size_t)ld * 4;
        int grid = 2 * num_sms();
        if (grid * LNB2_WARPS * 4 * R > prows) grid = cdiv(prows, LNB2_WARPS * 4 * R);
        int rows_per_warp = cdiv(prows, grid * LNB2_WARPS);
        rows_per_warp = (rows_per_warp + R - 1) / R * R;
        grid = cdiv(prows, rows_per_warp * LNB2_WARPS);
#define LNB2_CASE(NQ)                                                                                                   \
        do {                                                                                                            \
            cudaFuncSetAttribute(ln_bwd2_kernel<TDY, TM, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            ln_bwd2_kernel<TDY, TM, NQ><<<grid, LNB2_WARPS * 32, smem, s>>>((const TDY*)dy, x, gamma, mean, rstd, dres, dx,   \
                (TM*)dxm, drop_p, site, rng, dgamma, dbeta, B, L, d, halo, ld, rows_per_warp, stages, R);                 \
        } while (0)
        if (ld <= 128) LNB2_CASE(1); else if (ld <= 256) LNB2_CASE(2); else if (ld <= 384) LNB2_CASE(3);
        else LNB2_CASE(5);
#undef LNB2_CASE
        return;
    }
    int blocks = cdiv(B * L, 8 * 4);                 // 4 rows per warp (measured optimum of 2 / 4 / 16)
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
#define LNB_CASE(NP) ln_bwd_kernel<TDY, TM, NP><<<blocks, 256, 0, s>>>((const TDY*)dy, lddy, x, ldx, gamma, mean, rstd, \
        dres, lddres, dx, lddx, (TM*)dxm, lddxm, drop_p, site, rng, dgamma, dbeta, B, L, d, halo)
    if (d <= 192) LNB_CASE(3); else if (d <= 320) LNB_CASE(5); else if (d <= 576) LNB_CASE(9); else LNB_CASE(16);
#undef LNB_CASE
}

extern "C" int csi_layernorm_bwd(const void* dy, int lddy, int dy_dtype, const float* x, int ldx, const float* gamma,
                                 const float* mean, const float* rstd, const float* dres, int lddres, float* dx,
                                 int lddx, void* dxm, int lddxm, int dxm_dtype, float drop_p, unsigned drop_site,
                                 const unsigned long long* rng, float* dgamma, float* dbeta, int B, int L, int d,
                                 int halo, void* stream) {
    CSI_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "null pointer");
    CSI_CHECK_ARG(d % 2 == 0 && d <= 1024, "d must be even and <= 1024");
    CSI_CHECK_ARG(!(dxm && drop_p > 0.f) || rng, "dropout needs rng");
    if (B * L == 0) return CSI_OK;
    const bool b_dy = dy_dtype == CSI_BF16, b_m = dxm_dtype == CSI_BF16;
#define GO(TDY, TM) ln_bwd_launch<TDY, TM>(dy, lddy, x, ldx, gamma, mean, rstd, dres, lddres, dx, lddx, dxm, lddxm, \
        drop_p, drop_site, rng, dgamma, dbeta, B, L, d, halo, ST(stream))
    if (b_dy && b_m) GO(bf16, bf16); else if (b_dy) GO(bf16, float); else if (b_m) GO(float, bf16); else GO(float, float);
#undef GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ column reductions
// thread = one 8-channel chunk x one row lane; a CTA covers CR_ROWS valid token rows; lanes are combined in smem and
// the CTA issues one atomic per column.
#define CR_ROWS 64
#define CR_THREADS 256

template <typename T, bool SQ>
__global__ void __launch_bounds__(CR_THREADS) colreduce_kernel(const T* __restrict__ A, int lda, int B, int L, int halo,
                                                               int ncols, int CH, int RL, int rows_per_cta,
                                                               float* __restrict__ out_f, double* __restrict__ out_d, csi_grp grp) {
    extern __shared__ float red[];                      // [RL][CH*8] (x2 when SQ)
    const int ch = threadIdx.x % CH, rl = threadIdx.x / CH;
    const int total = B * L;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(total, r0 + rows_per_cta);
    float a[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
    if (rl < RL) {
        // eight rows of loads in flight per thread (the loop is latency-bound with one), rows walked without divisions
        RowWalk w;
        w.init(min(r0 + rl, total - 1), L, halo);
        int r = r0 + rl;
        for (; r + 7 * RL < r1; r += 8 * RL) {
            Raw8<T> x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                x[u].load(A + w.row() * lda + ch * 8);
                w.advance(RL);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float v[8];
                x[u].get(v);
#pragma unroll
                for (int j = 0; j < 8; ++j) { a[j] += v[j]; if (SQ) q[j] += v[j] * v[j]; }
            }
        }
        for (; r < r1; r += RL) {
            float v[8];
            load8f<T>(A + w.row() * lda + ch * 8, v);
            w.advance(RL);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += v[j]; if (SQ) q[j] += v[j] * v[j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[(rl * CH + ch) * 8 + j] = a[j];
            if (SQ) red[((RL + rl) * CH + ch) * 8 + j] = q[j];
        }
    }
    __syncthreads();
    // every CTA ends with one atomic per column on the same addresses: start each CTA at a different column so that
    // concurrent CTAs do not queue up on one address at a time
    const int rot = (int)((blockIdx.x * 61u) % (unsigned)(CH * 8));
    for (int c0 = threadIdx.x; c0 < CH * 8; c0 += blockDim.x) {
        const int c = c0 + rot < CH * 8 ? c0 + rot : c0 + rot - CH * 8;
        if (c >= ncols) continue;
        float s = 0.f, s2 = 0.f;
        for (int i = 0; i < RL; ++i) { s += red[i * CH * 8 + c]; if (SQ) s2 += red[(RL + i) * CH * 8 + c]; }
        if (SQ) { atomicAdd(out_d + c, (double)s); atomicAdd(out_d + ncols + c, (double)s2); }
        else { const int cc = grp_to_compact(c, grp); if (cc >= 0) atomicAdd(out_f + cc, s); }
    }
}

template <bool SQ>
static int colreduce_launch(const void* A, int lda, int dtype, int B, int L, int halo, int ncols, float* out_f, double* out_d,
                            csi_grp grp, cudaStream_t s) {
    const int CH = (ncols + 7) / 8;
    if (CH > CR_THREADS) { csi_set_error("colreduce: more than 2048 columns"); return CSI_ERR_ARG; }
    const int RL = CR_THREADS / CH;
    const size_t smem = (size_t)RL * CH * 8 * sizeof(float) * (SQ ? 2 : 1);
    // at most two CTAs per SM, each walking one contiguous slab of rows: every CTA ends with one atomic per column on the
    // same ncols addresses, so fewer, longer CTAs mean fewer serialised same-address atomics
    int grid = cdiv(B * L, CR_ROWS);
    if (grid > 2 * num_sms()) grid = 2 * num_sms();
    const int rows_per_cta = cdiv(B * L, grid);
    grid = cdiv(B * L, rows_per_cta);
    if (dtype == CSI_BF16) colreduce_kernel<bf16, SQ><<<grid, CR_THREADS, smem, s>>>((const bf16*)A, lda, B, L, halo, ncols, CH, RL, rows_per_cta, out_f, out_d, grp);
    else colreduce_kernel<float, SQ><<<grid, CR_THREADS, smem, s>>>((const float*)A, lda, B, L, halo, ncols, CH, RL, rows_per_cta, out_f, out_d, grp);
    return CSI_OK;
}

extern "C" int csi_colsum_tokens(const void* A, int lda, int dtype, int B, int L, int halo, int ncols, csi_grp grp, float* out,
                                 void* stream) {
    CSI_CHECK_ARG(A && out, "null pointer");
    CSI_CHECK_ARG(lda % 8 == 0 && ((ncols + 7) & ~7) <= lda, "lda must be a multiple of 8 covering ncols rounded up to 8");
    if (B * L == 0 || ncols == 0) return CSI_OK;
    int rc = colreduce_launch<false>(A, lda, dtype, B, L, halo, ncols, out, nullptr, grp, ST(stream));
    if (rc) return rc;
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

extern "C" int csi_bn_stats(const void* z, int ldz, int dtype, int B, int L, int halo, int ncols, double* sums,
                            void* stream) {
    CSI_CHECK_ARG(z && sums, "null pointer");
    CSI_CHECK_ARG(ldz % 8 == 0 && ((ncols + 7) & ~7) <= ldz, "ldz must be a multiple of 8 covering ncols");
    if (B * L == 0) return CSI_OK;
    int rc = colreduce_launch<true>(z, ldz, dtype, B, L, halo, ncols, nullptr, sums, csi_grp{0, 0}, ST(stream));
    if (rc) return rc;
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
// ------------------------------------------------------------------------------------------------ batchnorm
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int Dp, int d, int nbr, long long count,
                                   csi_ptr3 conv_bias, csi_ptr3 run_mean, csi_ptr3 run_var, csi_ptr3 nbt,
                                   float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, nc = nbr * Dp;
    if (q >= nc) return;
    const int br = q / Dp, c = q % Dp;
    if (c >= d) { mean[q] = 0.f; invstd[q] = 0.f; return; }
    const double m = sums[q] / (double)count;
    double var = sums[nc + q] / (double)count - m * m;
    if (var < 0.0) var = 0.0;
    mean[q] = (float)m;
    invstd[q] = (float)(1.0 / sqrt(var + (double)eps));
    float* rm = (float*)run_mean.p[br];
    float* rv = (float*)run_var.p[br];
    const float* cb = (const float*)conv_bias.p[br];
    const double unb = count > 1 ? (double)count / (double)(count - 1) : 1.0;
    rm[c] = (1.f - momentum) * rm[c] + momentum * ((float)m + cb[c]);
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)(var * unb);
    if (c == 0) { long long* n = (long long*)nbt.p[br]; n[0] += 1; }
}

extern "C" int csi_bn_finalize(const double* sums, int Dp, int d, int nbr, long long count, csi_ptr3 conv_bias,
                               csi_ptr3 run_mean, csi_ptr3 run_var, csi_ptr3 num_batches, float momentum, float eps,
                               float* mean, float* invstd, void* stream) {
    CSI_CHECK_ARG(sums && mean && invstd && nbr >= 1 && nbr <= 3 && count > 0, "bad argument");
    bn_finalize_kernel<<<cdiv(nbr * Dp, 128), 128, 0, ST(stream)>>>(sums, Dp, d, nbr, count, conv_bias, run_mean,
                                                                     run_var, num_batches, momentum, eps, mean, invstd);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

__global__ void bn_eval_prepare_kernel(int Dp, int d, int nbr, csi_ptr3 conv_bias, csi_ptr3 run_mean,
                                       csi_ptr3 run_var, float eps, float* __restrict__ mean,
                                       float* __restrict__ invstd) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nbr * Dp) return;
    const int br = q / Dp, c = q % Dp;
    if (c >= d) { mean[q] = 0.f; invstd[q] = 0.f; return; }
    mean[q] = ((const float*)run_mean.p[br])[c] - ((const float*)conv_bias.p[br])[c];
    invstd[q] = rsqrtf(((const float*)run_var.p[br])[c] + eps);
}

extern "C" int csi_bn_eval_prepare(int Dp, int d, int nbr, csi_ptr3 conv_bias, csi_ptr3 run_mean, csi_ptr3 run_var,
                                   float eps, float* mean, float* invstd, void* stream) {
    CSI_CHECK_ARG(mean && invstd && nbr >= 1 && nbr <= 3, "bad argument");
    bn_eval_prepare_kernel<<<cdiv(nbr * Dp, 128), 128, 0, ST(stream)>>>(Dp, d, nbr, conv_bias, run_mean, run_var, eps,
                                                                         mean, invstd);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

struct DropCfg {
    float p_branch, p_out;
    unsigned site_branch, site_out;
};

// Thread = one 8-channel chunk (fixed for the thread's lifetime, so the per-channel affine terms stay in registers)
// x one row lane; a CTA walks BNA_ROWS valid token rows.
#define BNA_THREADS 256
#define BNA_ITERS 8
// resident CTAs per SM of the BatchNorm block kernels.  CSI_BN_CTAS=3 selects a build capped at 80 registers (21 instead
// of 14 warps per SM, the per-channel affine terms spill to local memory): measured 25-45 % SLOWER at B=256, F=270
// (fwd 299 -> 379 us, bwd 543 -> 796 us per step), so 2 stays the default; kept for A/B runs.
static int bn_ctas() {
    static int v = 0;
    if (!v) { const char* e = getenv("CSI_BN_CTAS"); v = (e && e[0] == '3') ? 3 : 2; }
    return v;
}

// y = zhat*gamma + beta = z*a + b with a = invstd*gamma, b = beta - mean*a (0 for pad channels)
__device__ __forceinline__ void bn_affine8(const float* __restrict__ mean, const float* __restrict__ invstd, const float* ga,
                                           const float* be, int q0, int c0, int d, float (&a)[8], float (&b)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool ok = (c0 + j) < d;
        const float is = ok ? invstd[q0 + j] : 0.f, g = ok ? ga[c0 + j] : 0.f;
        a[j] = is * g;
        b[j] = ok ? be[c0 + j] - mean[q0 + j] * a[j] : 0.f;
    }
}

template <typename T, int MINB>
__global__ void __launch_bounds__(BNA_THREADS, MINB) bn_act_fwd_kernel(
    const T* __restrict__ z, int ldz, const float* __restrict__ mean, const float* __restrict__ invstd, csi_ptr3 gamma,
    csi_ptr3 beta, const float* __restrict__ t_res, int ldt, float* __restrict__ out, int ldo, int B, int L, int d, int Dp,
    int halo, int nbr, DropCfg dcfg, const unsigned long long* __restrict__ rng, int CH, int RL, int rows_per_cta,
    unsigned int* __restrict__ masks) {
    const int ch = threadIdx.x % CH, rl = threadIdx.x / CH;       // blockDim.x == RL * CH
    const int total = B * L, ld8 = Dp >> 3;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(total, r0 + rows_per_cta);
    float a[3][8], b[3][8];
#pragma unroll
    for (int br = 0; br < 3; ++br) {
        if (br < nbr) bn_affine8(mean, invstd, (const float*)gamma.p[br], (const float*)beta.p[br], br * Dp + ch * 8, ch * 8, d, a[br], b[br]);
    }
    const bool db = dcfg.p_branch > 0.f, dro = dcfg.p_out > 0.f;
    DropCtx cb, co;
    if (db) cb = drop_ctx(rng, dcfg.p_branch);
    if (dro) co = drop_ctx(rng, dcfg.p_out);
    const float inv = 1.0f / nbr;
    auto process = [&](size_t row, const Raw8<T> (&zc)[3], const Raw8<float>& tc) {
        const unsigned long long idx8 = (unsigned long long)row * ld8 + ch;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned int word = 0xffffffffu;                          // keep bits: byte br = branch br, byte 3 = output dropout
#pragma unroll
        for (int br = 0; br < 3; ++br) {
            if (br < nbr) {
                float zv[8];
                zc[br].get(zv);
                unsigned int kb = 0xffu;
                if (db) kb = drop_bits8(cb, dcfg.site_branch + br, idx8);
                word = (word & ~(0xffu << (8 * br))) | (kb << (8 * br));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float y = zv[j] * a[br][j] + b[br][j];
                    if (db) y = ((kb >> j) & 1u) ? y * cb.inv_keep : 0.f;
                    acc[j] += leaky(y);
                }
            }
        }
        float tv[8];
        tc.get(tv);
        unsigned int ko = 0xffu;
        if (dro) ko = drop_bits8(co, dcfg.site_out, idx8);
        word = (word & 0x00ffffffu) | (ko << 24);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[j] * inv;
            if (dro) v = ((ko >> j) & 1u) ? v * co.inv_keep : 0.f;
            tv[j] += v;
        }
        store8f<float>(out + row * ldo + ch * 8, tv);
        if (masks) masks[idx8] = word;
    };
    RowWalk w;
    w.init(min(r0 + rl, total - 1), L, halo);
    for (int r = r0 + rl; r < r1; r += 2 * RL) {
        const size_t rowA = w.row();
        w.advance(RL);
        const bool hasB = r + RL < r1;
        const size_t rowB = hasB ? w.row() : rowA;
        w.advance(RL);
        Raw8<T> zA[3], zB[3];
        Raw8<float> tA, tB;
#pragma unroll
        for (int br = 0; br < 3; ++br)
            if (br < nbr) zA[br].load(z + rowA * ldz + br * Dp + ch * 8);
        tA.load(t_res + rowA * ldt + ch * 8);
#pragma unroll
        for (int br = 0; br < 3; ++br)
            if (br < nbr) zB[br].load(z + rowB * ldz + br * Dp + ch * 8);
        tB.load(t_res + rowB * ldt + ch * 8);
        process(rowA, zA, tA);
        if (hasB) process(rowB, zB, tB);
    }
}

extern "C" int csi_bn_act_fwd(const void* z, int ldz, int dtype, const float* mean, const float* invstd,
                              csi_ptr3 gamma, csi_ptr3 beta, const float* t_res, int ldt, float* out, int ldo, int B,
                              int L, int d, int halo, int nbr, float p_branch, unsigned site_branch, float p_out,
                              unsigned site_out, const unsigned long long* rng, unsigned int* masks, void* stream) {
    CSI_CHECK_ARG(z && mean && invstd && t_res && out, "null pointer");
    CSI_CHECK_ARG(nbr >= 1 && nbr <= 3, "bad shape");
    CSI_CHECK_ARG(!(p_branch > 0.f || p_out > 0.f) || rng, "dropout needs rng");
    if (B * L == 0) return CSI_OK;
    const int Dp = (d + 15) & ~15, CH = Dp / 8;
    CSI_CHECK_ARG(CH <= BNA_THREADS && ldz % 8 == 0 && ldt % 4 == 0 && ldo % 4 == 0 && ldt >= Dp && ldo >= Dp, "bad leading dimension");
    const int RL = BNA_THREADS / CH, threads = RL * CH, total = B * L;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    const int ctas = bn_ctas();
    int grid = ctas * num_sms();                                   // resident CTAs only, one contiguous slab of rows each
    if (grid > cdiv(total, 2 * RL)) grid = cdiv(total, 2 * RL);
    const int rows_per_cta = cdiv(total, grid);
    grid = cdiv(total, rows_per_cta);
#define BNA_GO(T, MINB) bn_act_fwd_kernel<T, MINB><<<grid, threads, 0, ST(stream)>>>((const T*)z, ldz, mean, invstd, gamma, beta, t_res, \
        ldt, out, ldo, B, L, d, Dp, halo, nbr, dc, rng, CH, RL, rows_per_cta, masks)
    if (dtype == CSI_BF16) { if (ctas == 3) BNA_GO(bf16, 3); else BNA_GO(bf16, 2); }
    else { if (ctas == 3) BNA_GO(float, 3); else BNA_GO(float, 2); }
#undef BNA_GO
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// Backward: thread = (8-channel chunk, branch) x row lane.  The grid is two CTAs per SM, each walking one contiguous
// slab of token rows (no wave tail); every thread keeps two rows of loads in flight.
//   MODE 0: per-channel sums of dy and dy*zhat -> red (doubles, atomics)
//   MODE 1: dz = gamma*invstd*(dy - s1/n - zhat*s2/n), plus dgamma/dbeta written once by CTA 0
#define BNB_MAXT 224
#ifndef BNB_ROWS_REDUCE
#define BNB_ROWS_REDUCE 4
#endif
#ifndef BNB_ROWS_DZ
#define BNB_ROWS_DZ 4
#endif

// REGEN: no stored keep bits, the dropout decisions are regenerated with Philox (tests / callers without a mask buffer); the
// production path reads the bits bn_act_fwd stored and does not carry the generator state through the row loop.
template <typename T, int MODE, int MINB, bool REGEN, int NR>
__global__ void __launch_bounds__(BNB_MAXT, MINB) bn_act_bwd_kernel(
    const float* __restrict__ dout, int lddo, const T* __restrict__ z, int ldz, const float* __restrict__ mean,
    const float* __restrict__ invstd, csi_ptr3 gamma, csi_ptr3 beta, double* __restrict__ red, int B, int L, int d, int Dp,
    int halo, int nbr, DropCfg dcfg, const unsigned long long* __restrict__ rng, T* __restrict__ dz, int lddz,
    csi_ptr3 dgamma, csi_ptr3 dbeta, int CH, int RL, int rows_per_cta, const unsigned int* __restrict__ masks) {
    extern __shared__ float sred[];                               // MODE 0: [RL][CH*nbr][16]
    const int ncombo = CH * nbr;
    const int combo = threadIdx.x % ncombo, rl = threadIdx.x / ncombo;       // blockDim.x == RL * ncombo
    const int ch = combo % CH, br = combo / CH;
    const int total = B * L, ld8 = Dp >> 3, nc = nbr * Dp;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(total, r0 + rows_per_cta);
    const int q0 = br * Dp + ch * 8, c0 = ch * 8;
    // y = zh*ga + be with zh = (z - mu)*is, written as y = z*a + b (a = is*ga, b = be - mu*a).
    //   MODE 0 accumulates s1 = sum dy and s2z = sum dy*z; sum dy*zh = is*(s2z - mu*s1) is formed once at the end.
    //   MODE 1: dz = gi*dy - gk1 - zh*gk2 = a*dy - k1 - z*k2 with k1 = gk1 - mu*is*gk2, k2 = is*gk2.
    // Two (three) per-channel vectors of state instead of four (seven): registers for more rows of loads in flight.
    float a[8], b[8], k1[8], k2[8], s1[8], s2[8];
    {
        const float* gp = (const float*)gamma.p[br];
        const float* bp = (const float*)beta.p[br];
        const float invn = 1.0f / ((float)B * (float)L);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = (c0 + j) < d;
            const float mu = ok ? mean[q0 + j] : 0.f, is = ok ? invstd[q0 + j] : 0.f;
            a[j] = ok ? is * gp[c0 + j] : 0.f;
            b[j] = ok ? bp[c0 + j] - mu * a[j] : 0.f;
            s1[j] = s2[j] = 0.f;
            if (MODE == 1) {
                const float gk1 = ok ? a[j] * (float)red[q0 + j] * invn : 0.f;
                const float gk2 = ok ? a[j] * (float)red[nc + q0 + j] * invn : 0.f;
                k2[j] = is * gk2;
                k1[j] = gk1 - mu * k2[j];
            }
        }
    }
    const bool db = dcfg.p_branch > 0.f, dro = dcfg.p_out > 0.f;
    DropCtx cb, co;
    if (REGEN && db) cb = drop_ctx(rng, dcfg.p_branch);
    if (REGEN && dro) co = drop_ctx(rng, dcfg.p_out);
    const float inv = 1.0f / nbr;
    const float sc = inv * (dro ? 1.f / (1.f - dcfg.p_out) : 1.f) * (db ? 1.f / (1.f - dcfg.p_branch) : 1.f);
    auto process = [&](size_t row, const Raw8<float>& gn, const Raw8<T>& zn, unsigned int word) {
        const unsigned long long idx8 = (unsigned long long)row * ld8 + ch;
        float g[8], zv[8], o[8];
        gn.get(g);
        zn.get(zv);
        unsigned int kb = 0xffu, ko = 0xffu;                      // keep bits of the branch / output dropout
        if (!REGEN) { if (masks) { kb = (word >> (8 * br)) & 0xffu; ko = word >> 24; } }
        else {
            if (db) kb = drop_bits8(cb, dcfg.site_branch + br, idx8);
            if (dro) ko = drop_bits8(co, dcfg.site_out, idx8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaf(zv[j], a[j], b[j]);               // sign(y) is not changed by the (non-negative) keep scale
            const bool keep = ((kb >> j) & (ko >> j) & 1u) != 0u;
            const float dy = keep ? g[j] * sc * leaky_grad(y) : 0.f;
            if (MODE == 0) { s1[j] += dy; s2[j] = fmaf(dy, zv[j], s2[j]); }
            else o[j] = fmaf(-zv[j], k2[j], fmaf(a[j], dy, -k1[j]));
        }
        if (MODE == 1) store8f<T>(dz + row * lddz + q0, o);
    };
    // NR rows of loads in flight per thread: both passes are bound by load latency (ncu, reduce pass with two rows: 49 %
    // long-scoreboard stalls at 52 % issue utilisation)
    RowWalk w;
    w.init(min(r0 + rl, total - 1), L, halo);
    for (int r = r0 + rl; r < r1; r += NR * RL) {
        size_t rows[NR];
        bool has[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            has[i] = r + i * RL < r1;
            rows[i] = has[i] ? w.row() : rows[0];
            w.advance(RL);
        }
        Raw8<float> gR[NR];
        Raw8<T> zR[NR];
        unsigned int wR[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            gR[i].load(dout + rows[i] * lddo + c0);
            zR[i].load(z + rows[i] * ldz + q0);
            wR[i] = masks ? masks[rows[i] * ld8 + ch] : 0u;
        }
#pragma unroll
        for (int i = 0; i < NR; ++i)
            if (has[i]) process(rows[i], gR[i], zR[i], wR[i]);
    }
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = (c0 + j) < d;
            const float mu = ok ? mean[q0 + j] : 0.f, is = ok ? invstd[q0 + j] : 0.f;
            sred[((rl * ncombo + combo) * 2 + 0) * 8 + j] = s1[j];
            sred[((rl * ncombo + combo) * 2 + 1) * 8 + j] = is * (s2[j] - mu * s1[j]);       // sum dy * zhat of this thread's rows
        }
        __syncthreads();
        for (int e = threadIdx.x; e < ncombo * 16; e += blockDim.x) {
            const int cmb = e / 16, which = (e / 8) & 1, j = e & 7;
            const int cc = (cmb % CH) * 8 + j, bb = cmb / CH;
            if (cc >= d) continue;
            float s = 0.f;
            for (int i = 0; i < RL; ++i) s += sred[((i * ncombo + cmb) * 2 + which) * 8 + j];
            atomicAdd(red + which * nc + bb * Dp + cc, (double)s);
        }
    }
    if (MODE == 1 && blockIdx.x == 0 && rl == 0) {
        float* dg = (float*)dgamma.p[br];
        float* dbp = (float*)dbeta.p[br];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < d) { dg[c0 + j] += (float)red[nc + q0 + j]; dbp[c0 + j] += (float)red[q0 + j]; }
    }
}


template <int MODE>
static int bn_bwd_launch(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean, const float* invstd,
                         csi_ptr3 gamma, csi_ptr3 beta, double* red, int B, int L, int d, int halo, int nbr, DropCfg dc,
                         const unsigned long long* rng, void* dz, int lddz, csi_ptr3 dgamma, csi_ptr3 dbeta, const unsigned int* masks,
                         cudaStream_t s) {
    const int Dp = (d + 15) & ~15, CH = Dp / 8;
    if (CH * nbr > BNB_MAXT) { csi_set_error("bn_act_bwd: more than %d channels", BNB_MAXT * 8 / nbr); return CSI_ERR_ARG; }
    const int RL = BNB_MAXT / (CH * nbr), threads = RL * CH * nbr;
    const size_t smem = MODE == 0 ? (size_t)RL * CH * nbr * 16 * sizeof(float) : 0;
    const int ctas = bn_ctas();
    int grid = ctas * num_sms();
    const int total = B * L;
    if (grid > cdiv(total, 2 * RL)) grid = cdiv(total, 2 * RL);
    const int rows_per_cta = cdiv(total, grid);
    grid = cdiv(total, rows_per_cta);
    const bool regen = !masks && (dc.p_branch > 0.f || dc.p_out > 0.f);
    // rows of loads in flight per thread (A/B: CSI_BN_ROWS_REDUCE / CSI_BN_ROWS_DZ = 2 | 4 | 6); the Philox path keeps 2
    static int nr_env[2] = {0, 0};
    if (!nr_env[MODE]) {
        const char* e = getenv(MODE == 0 ? "CSI_BN_ROWS_REDUCE" : "CSI_BN_ROWS_DZ");
        const int v = e ? atoi(e) : 0;
        nr_env[MODE] = (v == 2 || v == 4 || v == 6) ? v : (MODE == 0 ? BNB_ROWS_REDUCE : BNB_ROWS_DZ);
    }
    const int nr = regen ? 2 : nr_env[MODE];
#define BNB_GO3(T, MINB, RG, NR) bn_act_bwd_kernel<T, MODE, MINB, RG, NR><<<grid, threads, smem, s>>>(dout, lddo, (const T*)z, ldz, mean, invstd, gamma, beta, \
        red, B, L, d, Dp, halo, nbr, dc, rng, (T*)dz, lddz, dgamma, dbeta, CH, RL, rows_per_cta, masks)
#define BNB_GO2(T, MINB, RG) do { if (nr == 6) BNB_GO3(T, MINB, RG, 6); else if (nr == 4) BNB_GO3(T, MINB, RG, 4); else BNB_GO3(T, MINB, RG, 2); } while (0)
#define BNB_GO(T, MINB) do { if (regen) BNB_GO3(T, MINB, true, 2); else BNB_GO2(T, MINB, false); } while (0)
    if (dtype == CSI_BF16) { if (ctas == 3) BNB_GO(bf16, 3); else BNB_GO(bf16, 2); }
    else { if (ctas == 3) BNB_GO(float, 3); else BNB_GO(float, 2); }
#undef BNB_GO2
#undef BNB_GO3
#undef BNB_GO
    return CSI_OK;
}

extern "C" int csi_bn_act_bwd_reduce(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                                     const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, int B, int L, int d, int halo,
                                     int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                                     const unsigned long long* rng, const unsigned int* masks, double* red, void* stream) {
    CSI_CHECK_ARG(dout && z && mean && invstd && red, "null pointer");
    CSI_CHECK_ARG(nbr >= 1 && nbr <= 3, "bad shape");
    if (B * L == 0) return CSI_OK;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    csi_ptr3 none{{nullptr, nullptr, nullptr}};
    int rc = bn_bwd_launch<0>(dout, lddo, z, ldz, dtype, mean, invstd, gamma, beta, red, B, L, d, halo, nbr, dc, rng, nullptr, 0,
                              none, none, masks, ST(stream));
    if (rc) return rc;
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

extern "C" int csi_bn_act_bwd_dz(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                                 const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, const double* red, int B, int L, int d,
                                 int halo, int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                                 const unsigned long long* rng, const unsigned int* masks, void* dz, int lddz, csi_ptr3 dgamma,
                                 csi_ptr3 dbeta, void* stream) {
    CSI_CHECK_ARG(dout && z && mean && invstd && red && dz, "null pointer");
    CSI_CHECK_ARG(nbr >= 1 && nbr <= 3, "bad shape");
    if (B * L == 0) return CSI_OK;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    int rc = bn_bwd_launch<1>(dout, lddo, z, ldz, dtype, mean, invstd, gamma, beta, const_cast<double*>(red), B, L, d, halo, nbr,
                              dc, rng, dz, lddz, dgamma, dbeta, masks, ST(stream));
    if (rc) return rc;
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
// ------------------------------------------------------------------------------------------------ heads
// feat[b, n] = sum_{t <= L-k(n)} leaky(p[row(b,t), n]); block = (sample, 32 columns) x 8 row lanes
template <typename T>
__global__ void __launch_bounds__(256) head_reduce_fwd_kernel(const T* __restrict__ p, int ldp, int L, int halo, int N,
                                                              int n0, int k0, int k1, float* __restrict__ feat, int ldf) {
    __shared__ float part[8][33];
    const int b = blockIdx.x, n = blockIdx.y * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int Lp = L + 2 * halo;
    float a = 0.f;
    if (n < N) {
        const int tmax = L - (n < n0 ? k0 : k1);
        for (int t = rl; t <= tmax; t += 8) a += leaky(ldv<T>(p + ((size_t)b * Lp + halo + t) * ldp + n));
    }
    part[rl][threadIdx.x & 31] = a;
    __syncthreads();
    if (rl == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
        feat[(size_t)b * ldf + n] = s;
    }
}

// eight consecutive channels per thread (16-byte loads): block = one sample, thread = (8-channel chunk, row lane)
template <typename T>
__global__ void __launch_bounds__(256) head_reduce_fwd8_kernel(const T* __restrict__ p, int ldp, int L, int halo, int N,
                                                               int n0, int k0, int k1, float* __restrict__ feat, int ldf,
                                                               int NC, int RL) {
    extern __shared__ float hr_part[];                              // [RL][N]
    const int b = blockIdx.x, c = threadIdx.x % NC, rl = threadIdx.x / NC, n = c * 8;
    const int Lp = L + 2 * halo;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rl < RL) {
        const int tmax = L - (n < n0 ? k0 : k1);
        const T* src = p + ((size_t)b * Lp + halo + rl) * ldp + n;
        for (int t = rl; t <= tmax; t += RL, src += (size_t)RL * ldp) {
            float v[8];
            load8f<T>(src, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += leaky(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) hr_part[rl * N + n + j] = a[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < RL; ++r) s += hr_part[r * N + i];
        feat[(size_t)b * ldf + i] = s;
    }
}

extern "C" int csi_head_reduce_fwd(const void* p, int ldp, int dtype, int B, int L, int halo, int N, int n0, int k0,
                                   int k1, float* feat, int ldf, void* stream) {
    CSI_CHECK_ARG(p && feat, "null pointer");
    CSI_CHECK_ARG(k0 <= L && k1 <= L, "kernel longer than the sequence");
    if (B == 0) return CSI_OK;
    {
        const int es = dtype == CSI_BF16 ? 2 : 4, NC = N / 8;
        if (N % 8 == 0 && n0 % 8 == 0 && NC >= 1 && NC <= 256 && (ldp * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(p) % 16) == 0) {
            int RL = 256 / NC;
            if (RL > 32) RL = 32;
            const size_t smem = (size_t)RL * N * sizeof(float);
            if (dtype == CSI_BF16) head_reduce_fwd8_kernel<bf16><<<B, 256, smem, ST(stream)>>>((const bf16*)p, ldp, L, halo, N, n0, k0, k1, feat, ldf, NC, RL);
            else head_reduce_fwd8_kernel<float><<<B, 256, smem, ST(stream)>>>((const float*)p, ldp, L, halo, N, n0, k0, k1, feat, ldf, NC, RL);
            CSI_LAUNCH_CHECK();
            return CSI_OK;
        }
    }
    dim3 grid(B, cdiv(N, 32));
    if (dtype == CSI_BF16) head_reduce_fwd_kernel<bf16><<<grid, 256, 0, ST(stream)>>>((const bf16*)p, ldp, L, halo, N, n0, k0, k1, feat, ldf);
    else head_reduce_fwd_kernel<float><<<grid, 256, 0, ST(stream)>>>((const float*)p, ldp, L, halo, N, n0, k0, k1, feat, ldf);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

template <typename T>
__global__ void head_reduce_bwd_kernel(const float* __restrict__ dfeat, int ldf, const T* __restrict__ p, int ldp, int B,
                                       int L, int halo, int N, int n0, int k0, int k1, T* __restrict__ dp, int lddp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L * N) return;
    const int n = (int)(idx % N);
    const int r = (int)(idx / N);
    const int b = r / L, t = r % L, Lp = L + 2 * halo;
    const size_t row = (size_t)b * Lp + halo + t;
    float v = 0.f;
    if (t <= L - (n < n0 ? k0 : k1)) v = dfeat[(size_t)b * ldf + n] * leaky_grad(ldv<T>(p + row * ldp + n));
    stf<T>(dp + row * lddp + n, v);
}

// eight consecutive channels per thread (16-byte accesses; a chunk lies inside one conv because n0 % 8 == 0): the scalar form
// above moved 2 bytes per thread and ran at 0.1 of the HBM bandwidth (29-35 us at the head of each stream's backward)
template <typename T>
__global__ void __launch_bounds__(256) head_reduce_bwd8_kernel(const float* __restrict__ dfeat, int ldf, const T* __restrict__ p,
                                                               int ldp, int B, int L, int halo, int N, int n0, int k0, int k1,
                                                               T* __restrict__ dp, int lddp) {
    const int N8 = N >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L * N8) return;
    const int n = (int)(idx % N8) * 8;
    const int r = (int)(idx / N8);
    const int b = r / L, t = r % L, Lp = L + 2 * halo;
    const size_t row = (size_t)b * Lp + halo + t;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t <= L - (n < n0 ? k0 : k1)) {
        float pv[8], dv[8];
        load8f<T>(p + row * ldp + n, pv);
        load8f<float>(dfeat + (size_t)b * ldf + n, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = dv[j] * leaky_grad(pv[j]);
    }
    store8f<T>(dp + row * lddp + n, v);
}

extern "C" int csi_head_reduce_bwd(const float* dfeat, int ldf, const void* p, int ldp, int dtype, int B, int L,
                                   int halo, int N, int n0, int k0, int k1, void* dp, int lddp, void* stream) {
    CSI_CHECK_ARG(dfeat && p && dp, "null pointer");
    if (B == 0) return CSI_OK;
    const int es = dtype == CSI_BF16 ? 2 : 4;
    const bool vec = N % 8 == 0 && n0 % 8 == 0 && ldp % 8 == 0 && lddp % 8 == 0 && ldf % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(p) % 16) == 0 && (reinterpret_cast<uintptr_t>(dp) % 16) == 0 &&
                     (reinterpret_cast<uintptr_t>(dfeat) % 16) == 0 && (ldp * es) % 16 == 0 && (lddp * es) % 16 == 0;
    if (vec) {
        const long long n8 = (long long)B * L * (N / 8);
        if (dtype == CSI_BF16)
            head_reduce_bwd8_kernel<bf16><<<cdiv(n8, 256), 256, 0, ST(stream)>>>(dfeat, ldf, (const bf16*)p, ldp, B, L, halo, N, n0, k0, k1, (bf16*)dp, lddp);
        else
            head_reduce_bwd8_kernel<float><<<cdiv(n8, 256), 256, 0, ST(stream)>>>(dfeat, ldf, (const float*)p, ldp, B, L, halo, N, n0, k0, k1, (float*)dp, lddp);
        CSI_LAUNCH_CHECK();
        return CSI_OK;
    }
    const long long n = (long long)B * L * N;
    if (dtype == CSI_BF16)
        head_reduce_bwd_kernel<bf16><<<cdiv(n, 256), 256, 0, ST(stream)>>>(dfeat, ldf, (const bf16*)p, ldp, B, L, halo, N, n0, k0, k1, (bf16*)dp, lddp);
    else
        head_reduce_bwd_kernel<float><<<cdiv(n, 256), 256, 0, ST(stream)>>>(dfeat, ldf, (const float*)p, ldp, B, L, halo, N, n0, k0, k1, (float*)dp, lddp);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ dropout / cast
template <typename T>
__global__ void dropout_rows_kernel(const float* __restrict__ in, int ldi, T* __restrict__ out, int ldo, int rows,
                                    int cols, float p, unsigned site, const unsigned long long* __restrict__ rng) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    const int c = (int)(idx % cols), r = (int)(idx / cols);
    float v = in[(size_t)r * ldi + c];
    if (p > 0.f) {
        const DropCtx dc = drop_ctx(rng, p);
        v *= drop_scale1(dc, site, (unsigned long long)r, ((cols + 15) & ~15) >> 3, c);
    }
    stf<T>(out + (size_t)r * ldo + c, v);
}

extern "C" int csi_dropout_rows(const float* in, int ldi, void* out, int ldo, int out_dtype, int rows, int cols, float p,
                                unsigned site, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(in && out, "null pointer");
    CSI_CHECK_ARG(!(p > 0.f) || rng, "dropout needs rng");
    if (rows * cols == 0) return CSI_OK;
    const long long n = (long long)rows * cols;
    if (out_dtype == CSI_BF16) dropout_rows_kernel<bf16><<<cdiv(n, 256), 256, 0, ST(stream)>>>(in, ldi, (bf16*)out, ldo, rows, cols, p, site, rng);
    else dropout_rows_kernel<float><<<cdiv(n, 256), 256, 0, ST(stream)>>>(in, ldi, (float*)out, ldo, rows, cols, p, site, rng);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ BCE with logits
__global__ void __launch_bounds__(1024) bce_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y,
                                                   int ldy, int rows, int cols, float pw, float gscale,
                                                   float* __restrict__ loss, float* __restrict__ dz, int lddz) {
    __shared__ float red[32];
    const int n = rows * cols;
    const float invn = 1.0f / (float)n;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i / cols, c = i % cols;
        const float zv = z[(size_t)r * ldz + c], yv = y[(size_t)r * ldy + c];
        // log sigmoid(z) = min(z,0) - log1p(exp(-|z|))
        const float l1p = log1pf(expf(-fabsf(zv)));
        const float ls_pos = fminf(zv, 0.f) - l1p, ls_neg = fminf(-zv, 0.f) - l1p;
        acc -= pw * yv * ls_pos + (1.f - yv) * ls_neg;
        if (dz) {
            const float sg = 1.f / (1.f + expf(-zv));
            dz[(size_t)r * lddz + c] = (sg * (pw * yv + 1.f - yv) - pw * yv) * invn * gscale;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = v * invn;
    }
}

extern "C" int csi_bce_logits(const float* z, int ldz, const float* y, int ldy, int rows, int cols, float pos_weight,
                              float grad_scale, float* loss, float* dz, int lddz, void* stream) {
    CSI_CHECK_ARG(z && y && loss && rows > 0 && cols > 0, "bad argument");
    bce_kernel<<<1, 1024, 0, ST(stream)>>>(z, ldz, y, ldy, rows, cols, pos_weight, grad_scale, loss, dz, lddz);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ SmoothL1
// torch.nn.SmoothL1Loss(reduction="mean", beta): the loss of the count-prediction sibling head
// (model/that_count_pred.py:399, train.py:91-97):  l = 0.5 d^2 / beta if |d| < beta else |d| - 0.5 beta,  d = z - y
__global__ void __launch_bounds__(1024) smooth_l1_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y,
                                                         int ldy, int rows, int cols, float beta, float gscale,
                                                         float* __restrict__ loss, float* __restrict__ dz, int lddz) {
    __shared__ float red[32];
    const int n = rows * cols;
    const float invn = 1.0f / (float)n;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i / cols, c = i % cols;
        const float dv = z[(size_t)r * ldz + c] - y[(size_t)r * ldy + c], ad = fabsf(dv);
        const bool quad = ad < beta;
        acc += quad ? 0.5f * dv * dv / beta : ad - 0.5f * beta;
        if (dz) dz[(size_t)r * lddz + c] = (quad ? dv / beta : (dv > 0.f ? 1.f : -1.f)) * invn * gscale;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = v * invn;
    }
}

extern "C" int csi_smooth_l1(const float* z, int ldz, const float* y, int ldy, int rows, int cols, float beta,
                             float grad_scale, float* loss, float* dz, int lddz, void* stream) {
    CSI_CHECK_ARG(z && y && loss && rows > 0 && cols > 0 && beta > 0.f, "bad argument");
    smooth_l1_kernel<<<1, 1024, 0, ST(stream)>>>(z, ldz, y, ldy, rows, cols, beta, grad_scale, loss, dz, lddz);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ permutation matching
// PermutationMatchingLoss of the five-head sibling (model/that_multi_head.py:309-342).  One thread per sample: log-sum-exp
// of each head, the H x H cost matrix cost[h][t] = lse[h] - z[h][class of slot t], then every permutation of the heads in
// lexicographic order (= itertools.permutations order; the FIRST minimum wins, as the reference keeps the incumbent
// unless loss < best).  loss = mean over B*H of cost[perm[t]][t]; dz = (softmax - onehot) * grad_scale / (B*H).
#define PCE_MAXH 6
__global__ void __launch_bounds__(256) perm_ce_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y, int ldy,
                                                      int B, int H, int C, int cp, float gscale, float* __restrict__ loss,
                                                      float* __restrict__ dz, int lddz, int* __restrict__ best_out) {
    __shared__ float red[8];
    const float invn = 1.0f / ((float)B * (float)H);
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* zr = z + (size_t)b * ldz;
        const float* yr = y + (size_t)b * ldy;
        int cls[PCE_MAXH];
        float lse[PCE_MAXH], cost[PCE_MAXH][PCE_MAXH];
#pragma unroll
        for (int t = 0; t < PCE_MAXH; ++t) {
            cls[t] = 0;
            if (t < H) {                                          // torch.argmax: first maximum
                float bv = yr[t * C];
                for (int c = 1; c < C; ++c) { const float v = yr[t * C + c]; if (v > bv) { bv = v; cls[t] = c; } }
            }
        }
#pragma unroll
        for (int h = 0; h < PCE_MAXH; ++h) {
            lse[h] = 0.f;
            if (h < H) {
                float m = zr[h * cp];
                for (int c = 1; c < C; ++c) m = fmaxf(m, zr[h * cp + c]);
                float sum = 0.f;
                for (int c = 0; c < C; ++c) sum += __expf(zr[h * cp + c] - m);
                lse[h] = m + __logf(sum);
            }
#pragma unroll
            for (int t = 0; t < PCE_MAXH; ++t) cost[h][t] = (h < H && t < H) ? lse[h] - zr[h * cp + cls[t]] : 0.f;
        }
        int perm[PCE_MAXH], best[PCE_MAXH];
#pragma unroll
        for (int t = 0; t < PCE_MAXH; ++t) perm[t] = best[t] = t;
        float bestv = INFINITY;
        for (;;) {
            float tot = 0.f;
#pragma unroll
            for (int t = 0; t < PCE_MAXH; ++t)
                if (t < H) {
                    float cv = 0.f;
#pragma unroll
                    for (int h = 0; h < PCE_MAXH; ++h) cv = perm[t] == h ? cost[h][t] : cv;
                    tot += cv;
                }
            tot /= (float)H;
            if (tot < bestv) {
                bestv = tot;
#pragma unroll
                for (int t = 0; t < PCE_MAXH; ++t) best[t] = perm[t];
            }
            // next permutation in lexicographic order
            int i = H - 2;
            while (i >= 0 && perm[i] > perm[i + 1]) --i;
            if (i < 0) break;
            int j = H - 1;
            while (perm[j] < perm[i]) --j;
            int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
            for (int a = i + 1, e = H - 1; a < e; ++a, --e) { tmp = perm[a]; perm[a] = perm[e]; perm[e] = tmp; }
        }
        for (int t = 0; t < H; ++t) {
            const int h = best[t];
            acc += cost[h][t];
            if (best_out) best_out[b * H + t] = h;
            if (dz) {
                float* dr = dz + (size_t)b * lddz + h * cp;
                for (int c = 0; c < cp; ++c) {
                    float g = 0.f;
                    if (c < C) g = (__expf(zr[h * cp + c] - lse[h]) - (c == cls[t] ? 1.f : 0.f)) * invn * gscale;
                    dr[c] = g;
                }
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = v * invn;
    }
}

extern "C" int csi_perm_ce(const float* z, int ldz, const float* y, int ldy, int B, int heads, int classes, int cpitch,
                           float grad_scale, float* loss, float* dz, int lddz, int* best_perm, void* stream) {
    CSI_CHECK_ARG(z && y && loss && B > 0 && classes > 0, "bad argument");
    CSI_CHECK_ARG(heads >= 1 && heads <= PCE_MAXH && cpitch >= classes && ldz >= heads * cpitch && ldy >= heads * classes &&
                  (!dz || lddz >= heads * cpitch), "1..6 heads, head pitch >= classes");
    perm_ce_kernel<<<1, 256, 0, ST(stream)>>>(z, ldz, y, ldy, B, heads, classes, cpitch, grad_scale, loss, dz, lddz, best_perm);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n4, float lr,
                                                   float b1, float b2, float eps, float wd,
                                                   const long long* __restrict__ step, float gscale) {
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const double t = (double)step[0];
        const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
        s_c[0] = (float)((double)lr / bc1);
        s_c[1] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const float step_size = s_c[0], inv_sqrt_bc2 = s_c[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define ADAM1(c)                                                              \
        { float gg = gv.c * gscale + wd * pv.c;                               \
          mv.c = b1 * mv.c + (1.f - b1) * gg;                                 \
          vv.c = b2 * vv.c + (1.f - b2) * gg * gg;                            \
          pv.c -= step_size * mv.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps); }
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
}

extern "C" int csi_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, const long long* step, float grad_scale,
                             void* stream) {
    CSI_CHECK_ARG(p && g && m && v && step, "null pointer");
    CSI_CHECK_ARG(n % 4 == 0, "arena length must be a multiple of 4");
    if (n == 0) return CSI_OK;
    int blocks = cdiv(n / 4, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<blocks, 256, 0, ST(stream)>>>(p, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay, step, grad_scale);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

__global__ void advance_kernel(unsigned long long* rng, long long* step) {
    if (rng) rng[1] += 1ull;
    if (step) step[0] += 1;
}
extern "C" int csi_advance_counters(unsigned long long* rng, long long* step, void* stream) {
    advance_kernel<<<1, 1, 0, ST(stream)>>>(rng, step);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ weight re-layout
template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ params, T* __restrict__ packed,
                                                   const csi_pack_entry* __restrict__ table) {
    // blockIdx.x walks the outer index of the DESTINATION (n in mode 0, c in mode 1), threads the inner one, taps in a
    // loop: no integer division on the hot path, coalesced stores; the strided fp32 reads hit L2 (19.6 MB of weights)
    const csi_pack_entry e = table[blockIdx.y];
    const float* src = params + e.src_off;
    T* dst = packed + e.dst_off;
    if (e.mode == 0) {
        for (int n = blockIdx.x; n < e.N; n += gridDim.x) {
            const long long drow = (long long)grp_to_padded(n, e.gn) * e.ld;
            for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
                const int cp = grp_to_padded(c, e.gc);
                const float* sp = src + ((long long)n * e.C + c) * e.k;
                for (int j = 0; j < e.k; ++j) stf<T>(dst + drow + (long long)j * e.P + cp, sp[j]);
            }
        }
    } else {
        for (int c = blockIdx.x; c < e.C; c += gridDim.x) {
            const long long drow = (long long)grp_to_padded(c, e.gc) * e.ld;
            for (int n = threadIdx.x; n < e.N; n += blockDim.x) {
                const int np = grp_to_padded(n, e.gn);
                const float* sp = src + ((long long)n * e.C + c) * e.k;
                for (int j = 0; j < e.k; ++j) stf<T>(dst + drow + (long long)(e.seg_base + j) * e.P + np, sp[j]);
            }
        }
    }
}

extern "C" int csi_pack_weights(const float* params, void* packed, int dtype, const csi_pack_entry* table,
                                int n_entries, int max_elems, void* stream) {
    CSI_CHECK_ARG(params && packed && table, "null pointer");
    if (n_entries == 0) return CSI_OK;
    (void)max_elems;
    const int bx = 96;                                 // outer-index blocks per table entry
    dim3 grid(bx, n_entries);
    if (dtype == CSI_BF16) pack_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(params, (bf16*)packed, table);
    else pack_kernel<float><<<grid, 256, 0, ST(stream)>>>(params, (float*)packed, table);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ prediction rule
__global__ void predict_counts_kernel(const float* __restrict__ z, int ldz, int rows, int users, int classes, float thr,
                                      int* __restrict__ counts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* zr = z + (size_t)r * ldz;
    int* cr = counts + (size_t)r * classes;
    for (int c = 0; c < classes; ++c) cr[c] = 0;
    // sigmoid is monotone: arg-max of the probabilities = first arg-max of the logits (numpy argmax keeps the first)
    const float zthr = logf(thr / (1.f - thr));
    for (int u = 0; u < users; ++u) {
        int best = 0;
        float bv = zr[u * classes];
        for (int c = 1; c < classes; ++c) {
            const float v = zr[u * classes + c];
            if (v > bv) { bv = v; best = c; }
        }
        if (bv > zthr) cr[best] += 1;
    }
}

extern "C" int csi_predict_counts(const float* logits, int ldz, int rows, int users, int classes, float threshold, int* counts,
                                  void* stream) {
    CSI_CHECK_ARG(logits && counts && users > 0 && classes > 0 && threshold > 0.f && threshold < 1.f, "bad argument");
    if (rows == 0) return CSI_OK;
    predict_counts_kernel<<<cdiv(rows, 128), 128, 0, ST(stream)>>>(logits, ldz, rows, users, classes, threshold, counts);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

__global__ void fill_kernel(float* p, long long n, float v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
extern "C" int csi_fill_f32(float* p, long long n, float v, void* stream) {
    CSI_CHECK_ARG(p || n == 0, "null pointer");
    if (n == 0) return CSI_OK;
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_kernel<<<blocks, 256, 0, ST(stream)>>>(p, n, v);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
