// Shared device/host helpers for libcsi_that.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/csi_that.h"

typedef __nv_bfloat16 bf16;

void csi_set_error(const char* fmt, ...);

#define CSI_CHECK_ARG(cond, msg)                                                   \
    do { if (!(cond)) { csi_set_error("%s: %s", __func__, msg); return CSI_ERR_ARG; } } while (0)

#define CSI_LAUNCH_CHECK()                                                         \
    do { cudaError_t e__ = cudaGetLastError();                                     \
         if (e__ != cudaSuccess) { csi_set_error("%s: %s", __func__, cudaGetErrorString(e__)); \
                                   return CSI_ERR_CUDA; } } while (0)

#define CSI_CUDA(call)                                                             \
    do { cudaError_t e__ = (call);                                                 \
         if (e__ != cudaSuccess) { csi_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
                                   return CSI_ERR_CUDA; } } while (0)

#define LEAKY_SLOPE 0.01f

// ---------------------------------------------------------------- dtype access
template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldv<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// two consecutive elements (p must be 2-element aligned)
template <typename T> __device__ __forceinline__ float2 ld2(const T* p);
template <> __device__ __forceinline__ float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld2<bf16>(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void st2(T* p, float2 v);
template <> __device__ __forceinline__ void st2<float>(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
template <> __device__ __forceinline__ void st2<bf16>(bf16* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __float22bfloat162_rn(v);
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- counter-based RNG (Philox4x32, 7 rounds)
// Philox4x32-7 (Salmon et al., SC'11: 7 rounds already pass BigCrush); one call yields 128 bits = the dropout
// decisions of 8 consecutive channels (16 bits each), or 4 augmentation samples.
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

struct RngKey { uint2 key; uint32_t step; };

__device__ __forceinline__ RngKey rng_load(const unsigned long long* rng) {
    RngKey r;
    unsigned long long seed = rng[0], step = rng[1];
    r.key = make_uint2((uint32_t)seed ^ (uint32_t)(step >> 32) * 0x85EBCA6Bu, (uint32_t)(seed >> 32) ^ 0x1234567u);
    r.step = (uint32_t)step;
    return r;
}

// 128 random bits for group `idx` of stream `site` at the current step
__device__ __forceinline__ uint4 rng_group(const RngKey& k, uint32_t site, unsigned long long idx) {
    return philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), site, k.step), k.key);
}

// ---- dropout: element (row, col) of a site belongs to group row*ld8 + col/8 (ld8 = round_up(cols,16)/8) and owns the
//      16-bit field col%8 of that group's 128 bits; it is kept iff field >= round(p*65536).
struct DropCtx { RngKey k; uint32_t thr; float inv_keep; };
__device__ __forceinline__ DropCtx drop_ctx(const unsigned long long* rng, float p) {
    DropCtx d;
    d.k = rng_load(rng);
    d.thr = (uint32_t)(p * 65536.0f + 0.5f);
    d.inv_keep = 1.f / (1.f - p);
    return d;
}
__device__ __forceinline__ uint32_t field16(const uint4& g, int j) {
    const uint32_t w = (j >> 1) == 0 ? g.x : ((j >> 1) == 1 ? g.y : ((j >> 1) == 2 ? g.z : g.w));
    return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}
// keep-scales (0 or 1/(1-p)) of the 8 channels of group idx8
__device__ __forceinline__ void drop_scales8(const DropCtx& d, uint32_t site, unsigned long long idx8, float (&ks)[8]) {
    const uint4 g = rng_group(d.k, site, idx8);
#pragma unroll
    for (int j = 0; j < 8; ++j) ks[j] = field16(g, j) >= d.thr ? d.inv_keep : 0.f;
}
// keep bits (bit j = channel j of the group is kept) -- the same decisions as drop_scales8
__device__ __forceinline__ uint32_t drop_bits8(const DropCtx& d, uint32_t site, unsigned long long idx8) {
    const uint4 g = rng_group(d.k, site, idx8);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m |= (field16(g, j) >= d.thr ? 1u : 0u) << j;
    return m;
}
// keep-scale of one element
__device__ __forceinline__ float drop_scale1(const DropCtx& d, uint32_t site, unsigned long long row, int ld8, int col) {
    const uint4 g = rng_group(d.k, site, row * (unsigned long long)ld8 + (col >> 3));
    return field16(g, col & 7) >= d.thr ? d.inv_keep : 0.f;
}

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : LEAKY_SLOPE * v; }
__device__ __forceinline__ float leaky_grad(float v) { return v > 0.f ? 1.f : LEAKY_SLOPE; }

// head-padding index maps (csi_grp in csi_that.h)
__host__ __device__ __forceinline__ int grp_to_padded(int i, csi_grp g) { return g.pad ? (i / g.valid) * g.pad + i % g.valid : i; }
__host__ __device__ __forceinline__ int grp_to_compact(int ip, csi_grp g) {
    if (!g.pad) return ip;
    const int r = ip % g.pad;
    return r < g.valid ? (ip / g.pad) * g.valid + r : -1;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
