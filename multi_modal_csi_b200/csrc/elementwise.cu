// Bandwidth-bound kernels of the THAT train step (sm_100a): input pooling/augmentation, Gaussian range
// encoding, LayerNorm, BatchNorm+activation, head reductions, loss, Adam, weight re-layout.
// All of them are coalesced along the channel axis, vectorised by 2 (channel counts are even: multiples of 10)
// and reduce with warp shuffles; none synchronises with the host.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void csi_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* csi_last_error(void) { return g_err; }
extern "C" int csi_abi_version(void) { return 1; }
extern "C" int csi_device_arch(int device) {
    cudaDeviceProp p;
    CSI_CUDA(cudaGetDeviceProperties(&p, device));
    return p.major * 10 + p.minor;
}

#define ST(s) ((cudaStream_t)(s))
#define SITE_AUG 9001u
#define SITE_AUG_SCALE 9002u

// ------------------------------------------------------------------------------------------------ pool_dual
#define POOL_K 20
#define POOL_TL 16

template <bool AUG>
__global__ void __launch_bounds__(288) pool_dual_kernel(
    const float* __restrict__ x, const long long* __restrict__ offs, const int* __restrict__ lens, int T, int F,
    const float* __restrict__ pe, int ld_pe, float* __restrict__ left, int ld_left, float* __restrict__ right,
    int ld_right, int halo, const unsigned long long* __restrict__ rng) {
    extern __shared__ float tile[];                    // [POOL_TL][F + 1]
    const int b = blockIdx.y, l0 = blockIdx.x * POOL_TL, L = T / POOL_K;
    const int ntl = min(POOL_TL, L - l0);
    const int Lp_l = L + 2 * halo, Lp_r = F + 2 * halo, FS = F + 1;
    const float* xb;
    int pad = 0;
    if (offs) { xb = x + offs[b]; pad = T - lens[b]; } else { xb = x + (size_t)b * T * F; }
    RngKey rk;
    float scale = 1.f;
    uint32_t keep_thr = 0;
    if (AUG) {
        rk = rng_load(rng);
        uint4 g = rng_group(rk, SITE_AUG_SCALE, (unsigned long long)b);
        scale = (float)g.x * (0.2f / 4294967296.0f) + 0.9f;           // U[0.9, 1.1)
        keep_thr = drop_threshold(0.04f);                             // Bernoulli(0.96) keep
    }
    for (int fp = threadIdx.x; fp < F / 2; fp += blockDim.x) {
        for (int tl = 0; tl < ntl; ++tl) {
            const int l = l0 + tl;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll 5
            for (int i = 0; i < POOL_K; ++i) {
                const int tau = l * POOL_K + i;
                float2 v = make_float2(0.f, 0.f);
                if (tau >= pad) v = *reinterpret_cast<const float2*>(xb + (size_t)(tau - pad) * F + 2 * fp);
                if (AUG) {
                    unsigned long long e2 = (((unsigned long long)b * T + tau) * F + 2 * fp) >> 1;
                    uint4 g = rng_group(rk, SITE_AUG, e2);
                    float u1 = ((float)g.x + 1.0f) * (1.0f / 4294967296.0f);
                    float u2 = (float)g.y * (1.0f / 4294967296.0f);
                    float r = sqrtf(-2.0f * __logf(u1));
                    float sn, cs;
                    __sincosf(6.283185307179586f * u2, &sn, &cs);
                    v.x = (v.x + 0.1f * r * cs) * scale * (g.z >= keep_thr ? 1.f : 0.f);
                    v.y = (v.y + 0.1f * r * sn) * scale * (g.w >= keep_thr ? 1.f : 0.f);
                }
                acc.x += v.x; acc.y += v.y;
            }
            acc.x *= (1.0f / POOL_K); acc.y *= (1.0f / POOL_K);
            tile[tl * FS + 2 * fp] = acc.x;
            tile[tl * FS + 2 * fp + 1] = acc.y;
            if (pe) { acc.x += pe[l * ld_pe + 2 * fp]; acc.y += pe[l * ld_pe + 2 * fp + 1]; }
            *reinterpret_cast<float2*>(left + ((size_t)b * Lp_l + halo + l) * ld_left + 2 * fp) = acc;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < F * POOL_TL; idx += blockDim.x) {
        const int f = idx / POOL_TL, tl = idx % POOL_TL;
        if (tl < ntl) right[((size_t)b * Lp_r + halo + f) * ld_right + l0 + tl] = tile[tl * FS + f];
    }
}

extern "C" int csi_pool_dual(const float* x, const long long* offs, const int* lens, int B, int T, int F,
                             const float* pe, int ld_pe, float* left, int ld_left, float* right, int ld_right,
                             int halo, int augment, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(x && left && right, "null pointer");
    CSI_CHECK_ARG(T % POOL_K == 0 && F % 2 == 0 && F <= 4096, "T must be a multiple of 20, F even and <= 4096");
    CSI_CHECK_ARG((offs == nullptr) == (lens == nullptr), "offs and lens go together");
    CSI_CHECK_ARG(!augment || rng, "augmentation needs rng");
    if (B == 0) return CSI_OK;
    const int L = T / POOL_K;
    int threads = ((F / 2 + 31) / 32) * 32;
    if (threads > 288) threads = 288;
    dim3 grid(cdiv(L, POOL_TL), B);
    size_t smem = (size_t)POOL_TL * (F + 1) * sizeof(float);
    if (augment) {
        if (smem > 48 * 1024) CSI_CUDA(cudaFuncSetAttribute(pool_dual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pool_dual_kernel<true><<<grid, threads, smem, ST(stream)>>>(x, offs, lens, T, F, pe, ld_pe, left, ld_left,
                                                                    right, ld_right, halo, rng);
    } else {
        if (smem > 48 * 1024) CSI_CUDA(cudaFuncSetAttribute(pool_dual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pool_dual_kernel<false><<<grid, threads, smem, ST(stream)>>>(x, offs, lens, T, F, pe, ld_pe, left, ld_left,
                                                                     right, ld_right, halo, rng);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
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
    const int Lp = L + 2 * halo, np = d >> 1;
    const float inv_d = 1.0f / d;
    RngKey rk;
    uint32_t thr = 0;
    float inv_keep = 1.f;
    const bool drop = (dxm != nullptr) && drop_p > 0.f;
    if (drop) { rk = rng_load(rng); thr = drop_threshold(drop_p); inv_keep = 1.f / (1.f - drop_p); }
    float2 ag[NP], ab[NP], gm[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        ag[j] = ab[j] = make_float2(0.f, 0.f);
        const int p = lane + 32 * j;
        gm[j] = p < np ? *reinterpret_cast<const float2*>(gamma + 2 * p) : make_float2(0.f, 0.f);
    }
    for (int r = gw; r < B * L; r += nw) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        const float mu = mean[row], rs = rstd[row];
        float2 g[NP], xh[NP];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) {
                float2 dv = ld2<TDY>(dy + row * lddy + 2 * p);
                float2 xv = *reinterpret_cast<const float2*>(x + row * ldx + 2 * p);
                xh[j] = make_float2((xv.x - mu) * rs, (xv.y - mu) * rs);
                g[j] = make_float2(dv.x * gm[j].x, dv.y * gm[j].y);
                s1 += g[j].x + g[j].y;
                s2 += g[j].x * xh[j].x + g[j].y * xh[j].y;
                ag[j].x += dv.x * xh[j].x; ag[j].y += dv.y * xh[j].y;
                ab[j].x += dv.x; ab[j].y += dv.y;
            } else { g[j] = xh[j] = make_float2(0.f, 0.f); }
        }
        s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = lane + 32 * j;
            if (p < np) {
                float2 o = make_float2(rs * (g[j].x - s1 - xh[j].x * s2), rs * (g[j].y - s1 - xh[j].y * s2));
                if (dres) {
                    float2 rv = *reinterpret_cast<const float2*>(dres + row * lddres + 2 * p);
                    o.x += rv.x; o.y += rv.y;
                }
                *reinterpret_cast<float2*>(dx + row * lddx + 2 * p) = o;
                if (dxm) {
                    if (drop) {
                        float2 ks = drop_scale2(rk, drop_site, (unsigned long long)row * d + 2 * p, thr, inv_keep);
                        o.x *= ks.x; o.y *= ks.y;
                    }
                    st2<TM>(dxm + row * lddxm + 2 * p, o);
                }
            }
        }
    }
    // block reduction of the per-warp dgamma / dbeta partials, then one atomic per column per block
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

template <typename TDY, typename TM>
static void ln_bwd_launch(const void* dy, int lddy, const float* x, int ldx, const float* gamma, const float* mean,
                          const float* rstd, const float* dres, int lddres, float* dx, int lddx, void* dxm, int lddxm,
                          float drop_p, unsigned site, const unsigned long long* rng, float* dgamma, float* dbeta,
                          int B, int L, int d, int halo, cudaStream_t s) {
    int blocks = cdiv(B * L, 8 * 16);
    if (blocks > 148 * 4) blocks = 148 * 4;
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

// ------------------------------------------------------------------------------------------------ column sums
#define CS_ROWS 128
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ A, int lda, int B, int L, int halo, int ncols,
                              float* __restrict__ out) {
    const int cp = blockIdx.y * blockDim.x + threadIdx.x;
    if (2 * cp >= ncols) return;
    const int Lp = L + 2 * halo, total = B * L;
    const int r0 = blockIdx.x * CS_ROWS, r1 = min(total, r0 + CS_ROWS);
    float2 a = make_float2(0.f, 0.f);
    for (int r = r0; r < r1; ++r) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        float2 v = ld2<T>(A + row * lda + 2 * cp);
        a.x += v.x; a.y += v.y;
    }
    atomicAdd(out + 2 * cp, a.x);
    if (2 * cp + 1 < ncols) atomicAdd(out + 2 * cp + 1, a.y);
}

extern "C" int csi_colsum_tokens(const void* A, int lda, int dtype, int B, int L, int halo, int ncols, float* out,
                                 void* stream) {
    CSI_CHECK_ARG(A && out, "null pointer");
    CSI_CHECK_ARG(lda % 2 == 0 && ((ncols + 1) & ~1) <= lda, "lda must be even and cover ncols rounded up to 2");
    if (B * L == 0 || ncols == 0) return CSI_OK;
    dim3 grid(cdiv(B * L, CS_ROWS), cdiv((ncols + 1) / 2, 128));
    if (dtype == CSI_BF16) colsum_kernel<bf16><<<grid, 128, 0, ST(stream)>>>((const bf16*)A, lda, B, L, halo, ncols, out);
    else colsum_kernel<float><<<grid, 128, 0, ST(stream)>>>((const float*)A, lda, B, L, halo, ncols, out);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ batchnorm
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ z, int ldz, int B, int L, int halo, int ncols,
                                double* __restrict__ sums) {
    const int cp = blockIdx.y * blockDim.x + threadIdx.x;
    if (2 * cp >= ncols) return;
    const int Lp = L + 2 * halo, total = B * L;
    const int r0 = blockIdx.x * CS_ROWS, r1 = min(total, r0 + CS_ROWS);
    float2 a = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
    for (int r = r0; r < r1; ++r) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        float2 v = ld2<T>(z + row * ldz + 2 * cp);
        a.x += v.x; a.y += v.y;
        q.x += v.x * v.x; q.y += v.y * v.y;
    }
    atomicAdd(sums + 2 * cp, (double)a.x);
    atomicAdd(sums + 2 * cp + 1, (double)a.y);
    atomicAdd(sums + ncols + 2 * cp, (double)q.x);
    atomicAdd(sums + ncols + 2 * cp + 1, (double)q.y);
}

extern "C" int csi_bn_stats(const void* z, int ldz, int dtype, int B, int L, int halo, int ncols, double* sums,
                            void* stream) {
    CSI_CHECK_ARG(z && sums, "null pointer");
    CSI_CHECK_ARG(ncols % 2 == 0 && ldz % 2 == 0, "ncols and ldz must be even");
    if (B * L == 0) return CSI_OK;
    dim3 grid(cdiv(B * L, CS_ROWS), cdiv(ncols / 2, 128));
    if (dtype == CSI_BF16) bn_stats_kernel<bf16><<<grid, 128, 0, ST(stream)>>>((const bf16*)z, ldz, B, L, halo, ncols, sums);
    else bn_stats_kernel<float><<<grid, 128, 0, ST(stream)>>>((const float*)z, ldz, B, L, halo, ncols, sums);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

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

// gradient/forward helper of one branch value: returns activation a and d(a)/d(y) (incl. branch dropout scale)
__device__ __forceinline__ void bn_branch(float zv, float mu, float is, float ga, float be, float ks, float& zh,
                                          float& act, float& dact) {
    zh = (zv - mu) * is;
    const float y = (zh * ga + be) * ks;
    act = leaky(y);
    dact = leaky_grad(y) * ks;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(
    const T* __restrict__ z, int ldz, const float* __restrict__ mean, const float* __restrict__ invstd,
    csi_ptr3 gamma, csi_ptr3 beta, const float* __restrict__ t_res, int ldt, float* __restrict__ out, int ldo, int B,
    int L, int d, int Dp, int halo, int nbr, DropCfg dc, const unsigned long long* __restrict__ rng) {
    const int np = d >> 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L * np) return;
    const int cp = (int)(idx % np);
    const int r = (int)(idx / np);
    const int Lp = L + 2 * halo;
    const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
    const bool db = dc.p_branch > 0.f, dout = dc.p_out > 0.f;
    RngKey rk;
    if (db || dout) rk = rng_load(rng);
    const unsigned long long eidx = (unsigned long long)row * d + 2 * cp;
    float2 acc = make_float2(0.f, 0.f);
    for (int br = 0; br < nbr; ++br) {
        const int q = br * Dp + 2 * cp;
        float2 zv = ld2<T>(z + row * ldz + q);
        float2 ks = make_float2(1.f, 1.f);
        if (db) ks = drop_scale2(rk, dc.site_branch + br, eidx, drop_threshold(dc.p_branch), 1.f / (1.f - dc.p_branch));
        const float* ga = (const float*)gamma.p[br];
        const float* be = (const float*)beta.p[br];
        float zh, a, da;
        bn_branch(zv.x, mean[q], invstd[q], ga[2 * cp], be[2 * cp], ks.x, zh, a, da); acc.x += a;
        bn_branch(zv.y, mean[q + 1], invstd[q + 1], ga[2 * cp + 1], be[2 * cp + 1], ks.y, zh, a, da); acc.y += a;
    }
    const float inv = 1.0f / nbr;
    acc.x *= inv; acc.y *= inv;
    if (dout) {
        float2 ks = drop_scale2(rk, dc.site_out, eidx, drop_threshold(dc.p_out), 1.f / (1.f - dc.p_out));
        acc.x *= ks.x; acc.y *= ks.y;
    }
    float2 tv = *reinterpret_cast<const float2*>(t_res + row * ldt + 2 * cp);
    *reinterpret_cast<float2*>(out + row * ldo + 2 * cp) = make_float2(tv.x + acc.x, tv.y + acc.y);
}

extern "C" int csi_bn_act_fwd(const void* z, int ldz, int dtype, const float* mean, const float* invstd,
                              csi_ptr3 gamma, csi_ptr3 beta, const float* t_res, int ldt, float* out, int ldo, int B,
                              int L, int d, int halo, int nbr, float p_branch, unsigned site_branch, float p_out,
                              unsigned site_out, const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(z && mean && invstd && t_res && out, "null pointer");
    CSI_CHECK_ARG(d % 2 == 0 && nbr >= 1 && nbr <= 3, "bad shape");
    CSI_CHECK_ARG(!(p_branch > 0.f || p_out > 0.f) || rng, "dropout needs rng");
    if (B * L == 0) return CSI_OK;
    const int Dp = (d + 15) & ~15;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    const long long n = (long long)B * L * (d / 2);
    if (dtype == CSI_BF16)
        bn_act_fwd_kernel<bf16><<<cdiv(n, 256), 256, 0, ST(stream)>>>((const bf16*)z, ldz, mean, invstd, gamma, beta, t_res,
                                                                       ldt, out, ldo, B, L, d, Dp, halo, nbr, dc, rng);
    else
        bn_act_fwd_kernel<float><<<cdiv(n, 256), 256, 0, ST(stream)>>>((const float*)z, ldz, mean, invstd, gamma, beta, t_res,
                                                                        ldt, out, ldo, B, L, d, Dp, halo, nbr, dc, rng);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// per-channel sums of dy and dy*zhat (thread = channel pair, block = chunk of rows)
template <typename T>
__global__ void bn_act_bwd_reduce_kernel(const float* __restrict__ dout, int lddo, const T* __restrict__ z, int ldz,
                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                         csi_ptr3 gamma, csi_ptr3 beta, int B, int L, int d, int Dp, int halo, int nbr,
                                         DropCfg dc, const unsigned long long* __restrict__ rng,
                                         double* __restrict__ red) {
    const int cp = blockIdx.y * blockDim.x + threadIdx.x;
    if (2 * cp >= d) return;
    const int Lp = L + 2 * halo, total = B * L, nc = nbr * Dp;
    const int r0 = blockIdx.x * CS_ROWS, r1 = min(total, r0 + CS_ROWS);
    const bool db = dc.p_branch > 0.f, dro = dc.p_out > 0.f;
    RngKey rk;
    if (db || dro) rk = rng_load(rng);
    float2 s1[3], s2[3], mu[3], is[3], ga[3], be[3];
    for (int br = 0; br < nbr; ++br) {
        const int q = br * Dp + 2 * cp;
        s1[br] = s2[br] = make_float2(0.f, 0.f);
        mu[br] = make_float2(mean[q], mean[q + 1]);
        is[br] = make_float2(invstd[q], invstd[q + 1]);
        ga[br] = make_float2(((const float*)gamma.p[br])[2 * cp], ((const float*)gamma.p[br])[2 * cp + 1]);
        be[br] = make_float2(((const float*)beta.p[br])[2 * cp], ((const float*)beta.p[br])[2 * cp + 1]);
    }
    const float inv = 1.0f / nbr;
    for (int r = r0; r < r1; ++r) {
        const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
        const unsigned long long eidx = (unsigned long long)row * d + 2 * cp;
        float2 g = *reinterpret_cast<const float2*>(dout + row * lddo + 2 * cp);
        g.x *= inv; g.y *= inv;
        if (dro) {
            float2 ks = drop_scale2(rk, dc.site_out, eidx, drop_threshold(dc.p_out), 1.f / (1.f - dc.p_out));
            g.x *= ks.x; g.y *= ks.y;
        }
        for (int br = 0; br < nbr; ++br) {
            float2 zv = ld2<T>(z + row * ldz + br * Dp + 2 * cp);
            float2 ks = make_float2(1.f, 1.f);
            if (db) ks = drop_scale2(rk, dc.site_branch + br, eidx, drop_threshold(dc.p_branch), 1.f / (1.f - dc.p_branch));
            float zh, a, da;
            bn_branch(zv.x, mu[br].x, is[br].x, ga[br].x, be[br].x, ks.x, zh, a, da);
            s1[br].x += g.x * da; s2[br].x += g.x * da * zh;
            bn_branch(zv.y, mu[br].y, is[br].y, ga[br].y, be[br].y, ks.y, zh, a, da);
            s1[br].y += g.y * da; s2[br].y += g.y * da * zh;
        }
    }
    for (int br = 0; br < nbr; ++br) {
        const int q = br * Dp + 2 * cp;
        atomicAdd(red + q, (double)s1[br].x); atomicAdd(red + q + 1, (double)s1[br].y);
        atomicAdd(red + nc + q, (double)s2[br].x); atomicAdd(red + nc + q + 1, (double)s2[br].y);
    }
}

extern "C" int csi_bn_act_bwd_reduce(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                                     const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, int B, int L, int d, int halo,
                                     int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                                     const unsigned long long* rng, double* red, void* stream) {
    CSI_CHECK_ARG(dout && z && mean && invstd && red, "null pointer");
    CSI_CHECK_ARG(d % 2 == 0 && nbr >= 1 && nbr <= 3, "bad shape");
    if (B * L == 0) return CSI_OK;
    const int Dp = (d + 15) & ~15;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    dim3 grid(cdiv(B * L, CS_ROWS), cdiv(d / 2, 64));
    if (dtype == CSI_BF16)
        bn_act_bwd_reduce_kernel<bf16><<<grid, 64, 0, ST(stream)>>>(dout, lddo, (const bf16*)z, ldz, mean, invstd, gamma, beta,
                                                                     B, L, d, Dp, halo, nbr, dc, rng, red);
    else
        bn_act_bwd_reduce_kernel<float><<<grid, 64, 0, ST(stream)>>>(dout, lddo, (const float*)z, ldz, mean, invstd, gamma, beta,
                                                                      B, L, d, Dp, halo, nbr, dc, rng, red);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_dz_kernel(
    const float* __restrict__ dout, int lddo, const T* __restrict__ z, int ldz, const float* __restrict__ mean,
    const float* __restrict__ invstd, csi_ptr3 gamma, csi_ptr3 beta, const double* __restrict__ red, int B, int L, int d,
    int Dp, int halo, int nbr, DropCfg dc, const unsigned long long* __restrict__ rng, T* __restrict__ dz, int lddz,
    csi_ptr3 dgamma, csi_ptr3 dbeta) {
    const int np = d >> 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L * np) return;
    const int cp = (int)(idx % np);
    const int r = (int)(idx / np);
    const int Lp = L + 2 * halo, nc = nbr * Dp;
    const size_t row = (size_t)(r / L) * Lp + halo + (r % L);
    const bool db = dc.p_branch > 0.f, dro = dc.p_out > 0.f;
    RngKey rk;
    if (db || dro) rk = rng_load(rng);
    const unsigned long long eidx = (unsigned long long)row * d + 2 * cp;
    const float inv = 1.0f / nbr, invn = 1.0f / ((float)B * (float)L);
    float2 g = *reinterpret_cast<const float2*>(dout + row * lddo + 2 * cp);
    g.x *= inv; g.y *= inv;
    if (dro) {
        float2 ks = drop_scale2(rk, dc.site_out, eidx, drop_threshold(dc.p_out), 1.f / (1.f - dc.p_out));
        g.x *= ks.x; g.y *= ks.y;
    }
    for (int br = 0; br < nbr; ++br) {
        const int q = br * Dp + 2 * cp;
        float2 zv = ld2<T>(z + row * ldz + q);
        float2 ks = make_float2(1.f, 1.f);
        if (db) ks = drop_scale2(rk, dc.site_branch + br, eidx, drop_threshold(dc.p_branch), 1.f / (1.f - dc.p_branch));
        const float* ga = (const float*)gamma.p[br];
        const float* be = (const float*)beta.p[br];
        const float s1x = (float)red[q], s1y = (float)red[q + 1], s2x = (float)red[nc + q], s2y = (float)red[nc + q + 1];
        float zh, a, da;
        float2 o;
        bn_branch(zv.x, mean[q], invstd[q], ga[2 * cp], be[2 * cp], ks.x, zh, a, da);
        o.x = ga[2 * cp] * invstd[q] * (g.x * da - s1x * invn - zh * s2x * invn);
        bn_branch(zv.y, mean[q + 1], invstd[q + 1], ga[2 * cp + 1], be[2 * cp + 1], ks.y, zh, a, da);
        o.y = ga[2 * cp + 1] * invstd[q + 1] * (g.y * da - s1y * invn - zh * s2y * invn);
        st2<T>(dz + row * lddz + q, o);
        if (r == 0) {
            float* dg = (float*)dgamma.p[br];
            float* dbp = (float*)dbeta.p[br];
            dg[2 * cp] += s2x; dg[2 * cp + 1] += s2y;
            dbp[2 * cp] += s1x; dbp[2 * cp + 1] += s1y;
        }
    }
}

extern "C" int csi_bn_act_bwd_dz(const float* dout, int lddo, const void* z, int ldz, int dtype, const float* mean,
                                 const float* invstd, csi_ptr3 gamma, csi_ptr3 beta, const double* red, int B, int L, int d,
                                 int halo, int nbr, float p_branch, unsigned site_branch, float p_out, unsigned site_out,
                                 const unsigned long long* rng, void* dz, int lddz, csi_ptr3 dgamma, csi_ptr3 dbeta,
                                 void* stream) {
    CSI_CHECK_ARG(dout && z && mean && invstd && red && dz, "null pointer");
    CSI_CHECK_ARG(d % 2 == 0 && nbr >= 1 && nbr <= 3, "bad shape");
    if (B * L == 0) return CSI_OK;
    const int Dp = (d + 15) & ~15;
    DropCfg dc{p_branch, p_out, site_branch, site_out};
    const long long n = (long long)B * L * (d / 2);
    if (dtype == CSI_BF16)
        bn_act_bwd_dz_kernel<bf16><<<cdiv(n, 256), 256, 0, ST(stream)>>>(dout, lddo, (const bf16*)z, ldz, mean, invstd, gamma,
                                                                          beta, red, B, L, d, Dp, halo, nbr, dc, rng, (bf16*)dz,
                                                                          lddz, dgamma, dbeta);
    else
        bn_act_bwd_dz_kernel<float><<<cdiv(n, 256), 256, 0, ST(stream)>>>(dout, lddo, (const float*)z, ldz, mean, invstd, gamma,
                                                                           beta, red, B, L, d, Dp, halo, nbr, dc, rng,
                                                                           (float*)dz, lddz, dgamma, dbeta);
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

extern "C" int csi_head_reduce_fwd(const void* p, int ldp, int dtype, int B, int L, int halo, int N, int n0, int k0,
                                   int k1, float* feat, int ldf, void* stream) {
    CSI_CHECK_ARG(p && feat, "null pointer");
    CSI_CHECK_ARG(k0 <= L && k1 <= L, "kernel longer than the sequence");
    if (B == 0) return CSI_OK;
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

extern "C" int csi_head_reduce_bwd(const float* dfeat, int ldf, const void* p, int ldp, int dtype, int B, int L,
                                   int halo, int N, int n0, int k0, int k1, void* dp, int lddp, void* stream) {
    CSI_CHECK_ARG(dfeat && p && dp, "null pointer");
    if (B == 0) return CSI_OK;
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
        RngKey rk = rng_load(rng);
        v *= drop_scale(rk, site, (unsigned long long)idx, drop_threshold(p), 1.f / (1.f - p));
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
__global__ void pack_kernel(const float* __restrict__ params, T* __restrict__ packed,
                            const csi_pack_entry* __restrict__ table) {
    const csi_pack_entry e = table[blockIdx.y];
    const int total = e.N * e.C * e.k;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int j = i % e.k, c = (i / e.k) % e.C, n = i / (e.k * e.C);
        const float v = params[e.src_off + i];
        const long long dst = e.mode == 0 ? (long long)n * e.ld + (long long)j * e.P + c
                                          : (long long)c * e.ld + (long long)(e.seg_base + j) * e.P + n;
        stf<T>(packed + e.dst_off + dst, v);
    }
}

extern "C" int csi_pack_weights(const float* params, void* packed, int dtype, const csi_pack_entry* table,
                                int n_entries, int max_elems, void* stream) {
    CSI_CHECK_ARG(params && packed && table, "null pointer");
    if (n_entries == 0) return CSI_OK;
    int bx = cdiv(max_elems, 256);
    if (bx > 64) bx = 64;
    dim3 grid(bx, n_entries);
    if (dtype == CSI_BF16) pack_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(params, (bf16*)packed, table);
    else pack_kernel<float><<<grid, 256, 0, ST(stream)>>>(params, (float*)packed, table);
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
