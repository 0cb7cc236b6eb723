// CSI-as-image path (BASELINE config 4; reference benchmark/wifi_csi/model/cnn_2d.py:23-99):
//   BatchNorm2d -> Conv2d(k, stride s, no padding) -> LeakyReLU -> Dropout(0.2), three times, then BatchNorm2d, mean over
//   the image, Linear.
// Activations are NHWC matrices [B*H*W, C].  A strided Conv2d is (BatchNorm-apply fused) im2col -> the tcgen05 NT GEMM of
// gemm_tc3.cu (K ordered (kh, kw, c): every kernel row of a patch is ONE contiguous run of k*C input elements); its data
// gradient is the same GEMM against the transposed weights followed by a col2im GATHER (no atomics: an input pixel sums
// the ceil(k/s)^2 patches that cover it), its weight gradient the tcgen05 TN GEMM over (dZ, col).  Everything else on the
// path is the bandwidth-bound kernels below: 16/32-byte vector accesses, per-thread channel ownership so that column
// statistics need no shuffles, fp64 atomics only for the per-block partial sums.
#include "common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define C2_THREADS 256

// ---- 8 consecutive elements (32 B of fp32 / 16 B of bf16; p must be that aligned)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

static inline int c2_grid(long long items) {
    long long b = (items + C2_THREADS - 1) / C2_THREADS;
    const long long cap = 148LL * 8;                        // 8 resident CTAs of 256 threads per SM
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool c2_chan_ok(int C) { return C == 1 || (C % 8 == 0 && C <= 256 && 2048 % C == 0); }

// ------------------------------------------------------------------------------------------------ per-channel sums
// x: [rows, C] contiguous.  sums[c] += sum x, sums[C + c] += sum x^2 (fp64).  A thread's 8-element vectors always cover the
// same 8 channels (the grid stride in elements is a multiple of 2048, and 2048 % C == 0), so it accumulates in registers.
// MODE 0: plain statistics.  MODE 1 (BatchNorm backward): x is the BN input, g * g_scale the upstream gradient (row r of g = row
// r / g_div: the gradient of a spatial mean is shared by the g_div positions it averages); sums[c] += sum g,
// sums[C + c] += sum g * (x - mean[c]) * invstd[c].
template <typename T, typename TG, int MODE>
__global__ void __launch_bounds__(C2_THREADS) nhwc_sums_kernel(const T* __restrict__ x, const TG* __restrict__ g, long long g_div,
                                                               float g_scale, long long n, int C, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, double* __restrict__ sums) {
    __shared__ double sh[2 * 256];
    for (int i = threadIdx.x; i < 2 * C; i += C2_THREADS) sh[i] = 0.0;
    __syncthreads();
    const long long nv = n >> 3, stride = (long long)gridDim.x * C2_THREADS;
    const long long v0 = (long long)blockIdx.x * C2_THREADS + threadIdx.x;
    const int c0 = C == 1 ? 0 : (int)((v0 * 8) % C);
    float mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mu[j] = (MODE == 1) ? mean[C == 1 ? 0 : c0 + j] : 0.f;
        is[j] = (MODE == 1) ? invstd[C == 1 ? 0 : c0 + j] : 1.f;
    }
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    long long v = v0;
    if (MODE == 0) {
        // four independent 16/32-byte loads in flight per thread before the first use
        for (; v + 3 * stride < nv; v += 4 * stride) {
            float a0[8], a1[8], a2[8], a3[8];
            load8<T>(x + v * 8, a0);
            load8<T>(x + (v + stride) * 8, a1);
            load8<T>(x + (v + 2 * stride) * 8, a2);
            load8<T>(x + (v + 3 * stride) * 8, a3);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += (a0[j] + a1[j]) + (a2[j] + a3[j]);
                q[j] = fmaf(a0[j], a0[j], fmaf(a1[j], a1[j], fmaf(a2[j], a2[j], fmaf(a3[j], a3[j], q[j]))));
            }
        }
    }
    for (; v < nv; v += stride) {
        float a[8];
        load8<T>(x + v * 8, a);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += a[j]; q[j] = fmaf(a[j], a[j], q[j]); }
        } else {
            float gg[8];
            if (g_div == 1) load8<TG>(g + v * 8, gg);
            else {
                const long long row = (v * 8) / C;
                load8<TG>(g + (row / g_div) * C + (v * 8 - row * C), gg);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { gg[j] *= g_scale; s[j] += gg[j]; q[j] = fmaf(gg[j], (a[j] - mu[j]) * is[j], q[j]); }
        }
    }
    if (C == 1) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { ts += s[j]; tq += q[j]; }
        if (blockIdx.x == 0 && threadIdx.x == 0)                          // scalar tail of a single-channel image
            for (long long i = nv << 3; i < n; ++i) {
                const float a = ldv<T>(x + i);
                if (MODE == 0) { ts += a; tq = fmaf(a, a, tq); }
                else { const float gg = ldv<TG>(g + i / g_div) * g_scale; ts += gg; tq = fmaf(gg, (a - mu[0]) * is[0], tq); }
            }
        ts = warp_sum(ts); tq = warp_sum(tq);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], (double)ts); atomicAdd(&sh[1], (double)tq); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&sh[c0 + j], (double)s[j]); atomicAdd(&sh[C + c0 + j], (double)q[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += C2_THREADS) atomicAdd(&sums[i], sh[i]);
}

extern "C" int csi_nhwc_stats(const void* x, int dtype, long long rows, int C, double* sums, void* stream) {
    CSI_CHECK_ARG(x && sums && rows >= 0, "bad argument");
    CSI_CHECK_ARG(c2_chan_ok(C), "C must be 1 or a multiple of 8 that divides 2048 (<= 256)");
    const long long n = rows * C;
    if (n == 0) return CSI_OK;
    const int grid = c2_grid(n >> 3);
    if (dtype == CSI_BF16)
        nhwc_sums_kernel<bf16, float, 0><<<grid, C2_THREADS, 0, ST(stream)>>>((const bf16*)x, nullptr, 1, 1.f, n, C, nullptr, nullptr, sums);
    else
        nhwc_sums_kernel<float, float, 0><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)x, nullptr, 1, 1.f, n, C, nullptr, nullptr, sums);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

extern "C" int csi_bn2d_bwd_reduce(const void* g, int g_dtype, long long g_div, float g_scale, const void* x, int x_dtype, long long rows, int C,
                                   const float* mean, const float* invstd, double* sums, void* stream) {
    CSI_CHECK_ARG(g && x && mean && invstd && sums && rows >= 0 && g_div >= 1, "bad argument");
    CSI_CHECK_ARG(c2_chan_ok(C) && C >= 8, "C must be a multiple of 8 that divides 2048 (<= 256)");
    const long long n = rows * C;
    if (n == 0) return CSI_OK;
    const int grid = c2_grid(n >> 3);
#define RED(T, TG) nhwc_sums_kernel<T, TG, 1><<<grid, C2_THREADS, 0, ST(stream)>>>((const T*)x, (const TG*)g, g_div, g_scale, n, C, mean, invstd, sums)
    if (x_dtype == CSI_BF16) { if (g_dtype == CSI_BF16) RED(bf16, bf16); else RED(bf16, float); }
    else { if (g_dtype == CSI_BF16) RED(float, bf16); else RED(float, float); }
#undef RED
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ BatchNorm2d finalize
// train: batch mean / biased variance from the fp64 sums, running statistics updated with the unbiased variance
// (torch.nn.BatchNorm2d, momentum 0.1); eval: running statistics.  scale = gamma * invstd, shift = beta - mean * scale.
__global__ void bn2d_finalize_kernel(const double* __restrict__ sums, int C, long long count, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* run_mean, float* run_var, long long* nbt,
                                     float momentum, float eps, int training, float* mean, float* invstd, float* scale,
                                     float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float m, is;
    if (training) {
        const double mm = sums[c] / (double)count;
        double var = sums[C + c] / (double)count - mm * mm;
        if (var < 0.0) var = 0.0;
        m = (float)mm;
        is = (float)(1.0 / sqrt(var + (double)eps));
        if (run_mean) {
            const double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
            run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m;
            run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
            if (c == 0 && nbt) nbt[0] += 1;
        }
    } else {
        m = run_mean[c];
        is = rsqrtf(run_var[c] + eps);
    }
    mean[c] = m;
    invstd[c] = is;
    const float sc = gamma[c] * is;
    scale[c] = sc;
    shift[c] = beta[c] - m * sc;
}

extern "C" int csi_bn2d_finalize(const double* sums, int C, long long count, const float* gamma, const float* beta, float* run_mean,
                                 float* run_var, long long* nbt, float momentum, float eps, int training, float* mean,
                                 float* invstd, float* scale, float* shift, void* stream) {
    CSI_CHECK_ARG(gamma && beta && mean && invstd && scale && shift && C >= 1 && count >= 1, "bad argument");
    CSI_CHECK_ARG(training ? sums != nullptr : (run_mean && run_var), "train mode needs sums, eval mode running statistics");
    bn2d_finalize_kernel<<<cdiv(C, 128), 128, 0, ST(stream)>>>(sums, C, count, gamma, beta, run_mean, run_var, nbt, momentum, eps,
                                                                training, mean, invstd, scale, shift);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ im2col (+ BatchNorm apply)
// col[m, (kh*k + kw)*C + c] = x[b, oh*s + kh, ow*s + kw, c] * scale[c] + shift[c],  m = (b*OH + oh)*OW + ow; columns
// [k*k*C, Kp) are zero.  C >= 8: one thread moves 8 channels (one 16/32-byte access each way); C == 1 (the raw CSI image,
// fp32): one thread writes two neighbouring columns.
template <typename T>
__global__ void __launch_bounds__(C2_THREADS) im2col_vec_kernel(const T* __restrict__ x, int H, int W, int C, int k, int s, int OH,
                                                                int OW, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, T* __restrict__ col, int Kp,
                                                                long long M) {
    const int K = k * k * C, cpr = Kp >> 3, kc8 = (k * C) >> 3;
    const long long items = M * cpr;
    for (long long it = (long long)blockIdx.x * C2_THREADS + threadIdx.x; it < items; it += (long long)gridDim.x * C2_THREADS) {
        // M < 2^31 patch rows: 64-bit only for the item index and the final addresses
        const unsigned m = (unsigned)(it / (unsigned)cpr);
        const int ch = (int)(it - (long long)m * cpr);
        float v[8];
        if (ch * 8 < K) {
            const int kh = ch / kc8, r = ch - kh * kc8;              // r: 8-element chunk inside the k*C run of this kernel row
            const unsigned t = m / (unsigned)OW;
            const int ow = (int)(m - t * (unsigned)OW);
            const long long b = t / (unsigned)OH;
            const int oh = (int)(t - (unsigned)b * (unsigned)OH);
            const T* src = x + ((b * H + (long long)oh * s + kh) * W + (long long)ow * s) * C + r * 8;
            load8<T>(src, v);
            const int c0 = (r * 8) % C;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], scale[c0 + j], shift[c0 + j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        store8<T>(col + (long long)m * Kp + ch * 8, v);
    }
}

template <typename TI, typename T>
__global__ void __launch_bounds__(C2_THREADS) im2col_c1_kernel(const TI* __restrict__ x, int H, int W, int k, int s, int OH, int OW,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               T* __restrict__ col, int Kp, long long M) {
    const int K = k * k, cpr = Kp >> 1;
    const float sc = scale[0], sh = shift[0];
    const long long items = M * cpr;
    for (long long it = (long long)blockIdx.x * C2_THREADS + threadIdx.x; it < items; it += (long long)gridDim.x * C2_THREADS) {
        const long long m = it / cpr;
        const int j0 = (int)(it - m * cpr) * 2;
        const int ow = (int)(m % OW);
        const long long t = m / OW;
        const int oh = (int)(t % OH);
        const long long b = t / OH;
        const TI* base = x + (b * H + (long long)oh * s) * W + (long long)ow * s;
        float2 v = make_float2(0.f, 0.f);
        if (j0 < K) { const int kh = j0 / k, kw = j0 - kh * k; v.x = fmaf(ldv<TI>(base + (long long)kh * W + kw), sc, sh); }
        if (j0 + 1 < K) { const int kh = (j0 + 1) / k, kw = (j0 + 1) - kh * k; v.y = fmaf(ldv<TI>(base + (long long)kh * W + kw), sc, sh); }
        st2<T>(col + m * Kp + j0, v);
    }
}

// Single-channel image, compile-time kernel size / stride (conv 0 is 27x27 / 7): one CTA per (sample, output row).  The K
// input rows the output row reads are ONE contiguous run of K*W floats: they are normalised into shared memory once
// (each input pixel is fetched K/S = 3.9 times overall instead of once per patch that covers it), then the OW patch rows
// are written as whole 16/32-byte chunks.
template <typename T, int K, int S>
__global__ void __launch_bounds__(C2_THREADS) im2col_c1_tile_kernel(const float* __restrict__ x, int H, int W, int OH, int OW,
                                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                                    T* __restrict__ col, int Kp) {
    extern __shared__ float c1_tile[];                                  // [K][W]
    const int boh = blockIdx.x, b = boh / OH, oh = boh - b * OH;
    const float sc = scale[0], sh = shift[0];
    const float* src = x + ((long long)b * H + (long long)oh * S) * W;
    const int n = K * W;
    if ((reinterpret_cast<uintptr_t>(src) & 7) == 0 && (n & 1) == 0) {
        for (int i = threadIdx.x; i < (n >> 1); i += C2_THREADS) {
            const float2 v = reinterpret_cast<const float2*>(src)[i];
            c1_tile[2 * i] = fmaf(v.x, sc, sh);
            c1_tile[2 * i + 1] = fmaf(v.y, sc, sh);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += C2_THREADS) c1_tile[i] = fmaf(src[i], sc, sh);
    }
    __syncthreads();
    const int cpr = Kp >> 3;
    T* dst = col + (long long)boh * OW * Kp;
    for (int c = threadIdx.x; c < OW * cpr; c += C2_THREADS) {
        const int ow = c / cpr, j0 = (c - ow * cpr) * 8;
        const float* base = c1_tile + ow * S;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int jj = j0 + j, kh = jj / K, kw = jj - kh * K;
            v[j] = jj < K * K ? base[kh * W + kw] : 0.f;
        }
        store8<T>(dst + (long long)ow * Kp + j0, v);
    }
}

extern "C" int csi_im2col_bn(const void* x, int x_dtype, int B, int H, int W, int C, int k, int s, const float* scale,
                             const float* shift, void* col, int col_dtype, int Kp, void* stream) {
    CSI_CHECK_ARG(x && scale && shift && col && B >= 0 && k >= 1 && s >= 1 && H >= k && W >= k, "bad argument");
    CSI_CHECK_ARG(Kp % 16 == 0 && Kp >= k * k * C, "Kp must be a multiple of 16 holding k*k*C columns");
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
    const long long M = (long long)B * OH * OW;
    if (M == 0) return CSI_OK;
    if (C == 1 && k == 27 && s == 7 && x_dtype == CSI_F32 && (size_t)k * W * sizeof(float) <= 48 * 1024) {
        const size_t smem = (size_t)k * W * sizeof(float);
        if (col_dtype == CSI_BF16)
            im2col_c1_tile_kernel<bf16, 27, 7><<<B * OH, C2_THREADS, smem, ST(stream)>>>((const float*)x, H, W, OH, OW, scale, shift, (bf16*)col, Kp);
        else
            im2col_c1_tile_kernel<float, 27, 7><<<B * OH, C2_THREADS, smem, ST(stream)>>>((const float*)x, H, W, OH, OW, scale, shift, (float*)col, Kp);
    } else if (C == 1) {
        const int grid = c2_grid(M * (Kp >> 1));
        if (x_dtype == CSI_F32 && col_dtype == CSI_BF16)
            im2col_c1_kernel<float, bf16><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)x, H, W, k, s, OH, OW, scale, shift, (bf16*)col, Kp, M);
        else if (x_dtype == CSI_F32 && col_dtype == CSI_F32)
            im2col_c1_kernel<float, float><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)x, H, W, k, s, OH, OW, scale, shift, (float*)col, Kp, M);
        else { csi_set_error("csi_im2col_bn: the single-channel image is fp32"); return CSI_ERR_ARG; }
    } else {
        CSI_CHECK_ARG(C % 8 == 0 && x_dtype == col_dtype, "C must be a multiple of 8 and the patch matrix of the activation type");
        const int grid = c2_grid(M * (Kp >> 3));
        if (x_dtype == CSI_BF16)
            im2col_vec_kernel<bf16><<<grid, C2_THREADS, 0, ST(stream)>>>((const bf16*)x, H, W, C, k, s, OH, OW, scale, shift, (bf16*)col, Kp, M);
        else
            im2col_vec_kernel<float><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)x, H, W, C, k, s, OH, OW, scale, shift, (float*)col, Kp, M);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ col2im (gather)
// g[b, h, w, c] = sum over the patches covering the pixel of gcol[(b, oh, ow), (kh*k + kw)*C + c], kh = h - oh*s, kw = w - ow*s
template <typename T>
__global__ void __launch_bounds__(C2_THREADS) col2im_kernel(const T* __restrict__ gcol, int H, int W, int C, int k, int s, int OH,
                                                            int OW, int Kp, float* __restrict__ g, long long pixels) {
    const int c8n = C >> 3;
    const long long items = pixels * c8n;
    for (long long it = (long long)blockIdx.x * C2_THREADS + threadIdx.x; it < items; it += (long long)gridDim.x * C2_THREADS) {
        const long long px = it / c8n;
        const int c0 = (int)(it - px * c8n) * 8;
        const int w = (int)(px % W);
        const long long t = px / W;
        const int h = (int)(t % H);
        const long long b = t / H;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        // patches (oh, ow) with oh*s <= h < oh*s + k
        const int oh_hi = min(h / s, OH - 1), oh_lo = max(0, (h - k + s) / s);      // ceil((h - k + 1) / s)
        const int ow_hi = min(w / s, OW - 1), ow_lo = max(0, (w - k + s) / s);
        for (int oh = oh_lo; oh <= oh_hi; ++oh) {
            const int kh = h - oh * s;
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                const int kw = w - ow * s;
                float v[8];
                load8<T>(gcol + ((b * OH + oh) * OW + ow) * (long long)Kp + (kh * k + kw) * C + c0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[j];
            }
        }
        store8<float>(g + px * C + c0, acc);
    }
}

extern "C" int csi_col2im(const void* gcol, int dtype, int B, int H, int W, int C, int k, int s, int Kp, float* g, void* stream) {
    CSI_CHECK_ARG(gcol && g && C % 8 == 0 && k >= 1 && s >= 1 && H >= k && W >= k && Kp >= k * k * C, "bad argument");
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
    const long long pixels = (long long)B * H * W;
    if (pixels == 0) return CSI_OK;
    const int grid = c2_grid(pixels * (C >> 3));
    if (dtype == CSI_BF16) col2im_kernel<bf16><<<grid, C2_THREADS, 0, ST(stream)>>>((const bf16*)gcol, H, W, C, k, s, OH, OW, Kp, g, pixels);
    else col2im_kernel<float><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)gcol, H, W, C, k, s, OH, OW, Kp, g, pixels);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ LeakyReLU + Dropout
// y = dropout_p(leaky(z)); the keep bits of every 8-channel group are stored (1 byte) for the backward pass
template <typename T>
__global__ void __launch_bounds__(C2_THREADS) act_drop_fwd_kernel(const T* __restrict__ z, T* __restrict__ y, long long groups,
                                                                  float p, unsigned site, const unsigned long long* __restrict__ rng,
                                                                  unsigned char* __restrict__ mask) {
    DropCtx dc;
    const bool drop = p > 0.f;
    if (drop) dc = drop_ctx(rng, p);
    for (long long gi = (long long)blockIdx.x * C2_THREADS + threadIdx.x; gi < groups; gi += (long long)gridDim.x * C2_THREADS) {
        float v[8];
        load8<T>(z + gi * 8, v);
        uint32_t bits = 0xFFu;
        if (drop) { bits = drop_bits8(dc, site, (unsigned long long)gi); mask[gi] = (unsigned char)bits; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = leaky(v[j]);
            v[j] = drop ? (((bits >> j) & 1u) ? a * dc.inv_keep : 0.f) : a;
        }
        store8<T>(y + gi * 8, v);
    }
}

extern "C" int csi_act_drop_fwd(const void* z, void* y, int dtype, long long n, float p, unsigned site,
                                const unsigned long long* rng, unsigned char* mask, void* stream) {
    CSI_CHECK_ARG(z && y && n >= 0 && n % 8 == 0, "element count must be a multiple of 8");
    CSI_CHECK_ARG(!(p > 0.f) || (rng && mask), "dropout needs rng and a mask buffer");
    if (n == 0) return CSI_OK;
    const int grid = c2_grid(n >> 3);
    if (dtype == CSI_BF16) act_drop_fwd_kernel<bf16><<<grid, C2_THREADS, 0, ST(stream)>>>((const bf16*)z, (bf16*)y, n >> 3, p, site, rng, mask);
    else act_drop_fwd_kernel<float><<<grid, C2_THREADS, 0, ST(stream)>>>((const float*)z, (float*)y, n >> 3, p, site, rng, mask);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ BatchNorm2d backward (apply)
// x: BN input (= y of the previous block = dropout(leaky(zprev))), g: gradient w.r.t. the BN output (row r of g = r / g_div),
// sums: [sum g, sum g*xhat] per channel (csi_bn2d_bwd_reduce).
//   gx = gamma * invstd * (g - sum_g / n - xhat * sum_gxhat / n)
//   gz = gx * keep / (1 - p) * leaky'(zprev)                        (activation + dropout backward of the previous block)
// Block 0 also adds the affine gradients: dgamma += sum_gxhat, dbeta += sum_g.
template <typename T>
__global__ void __launch_bounds__(C2_THREADS) bn2d_bwd_apply_kernel(const float* __restrict__ g, long long g_div, float g_scale, const T* __restrict__ x,
                                                                    const T* __restrict__ zprev, const unsigned char* __restrict__ mask,
                                                                    float drop_p, long long n, int C, const float* __restrict__ mean,
                                                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                    const double* __restrict__ sums, double inv_count, T* __restrict__ gz,
                                                                    float* dgamma, float* dbeta) {
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += C2_THREADS) { dbeta[c] += (float)sums[c]; dgamma[c] += (float)sums[C + c]; }
    const long long nv = n >> 3, stride = (long long)gridDim.x * C2_THREADS;
    const long long v0 = (long long)blockIdx.x * C2_THREADS + threadIdx.x;
    const int c0 = (int)((v0 * 8) % C);
    float mu[8], is[8], a[8], m0[8], m1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j]; a[j] = gamma[c0 + j] * is[j];
        m0[j] = (float)(sums[c0 + j] * inv_count); m1[j] = (float)(sums[C + c0 + j] * inv_count);
    }
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    for (long long v = v0; v < nv; v += stride) {
        float xv[8], gg[8], zv[8];
        load8<T>(x + v * 8, xv);
        load8<T>(zprev + v * 8, zv);
        if (g_div == 1) load8<float>(g + v * 8, gg);
        else {
            const long long row = (v * 8) / C;
            load8<float>(g + (row / g_div) * C + (v * 8 - row * C), gg);
        }
        const uint32_t bits = drop_p > 0.f ? mask[v] : 0xFFu;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - mu[j]) * is[j];
            const float gx = a[j] * (gg[j] * g_scale - m0[j] - xh * m1[j]);
            gg[j] = ((bits >> j) & 1u) ? gx * inv_keep * leaky_grad(zv[j]) : 0.f;
        }
        store8<T>(gz + v * 8, gg);
    }
}

extern "C" int csi_bn2d_bwd_apply(const float* g, long long g_div, float g_scale, const void* x, const void* zprev, int dtype, const unsigned char* mask,
                                  float drop_p, long long rows, int C, const float* mean, const float* invstd, const float* gamma,
                                  const double* sums, void* gz, float* dgamma, float* dbeta, void* stream) {
    CSI_CHECK_ARG(g && x && zprev && mean && invstd && gamma && sums && gz && dgamma && dbeta && rows >= 1 && g_div >= 1, "bad argument");
    CSI_CHECK_ARG(c2_chan_ok(C) && C >= 8, "C must be a multiple of 8 that divides 2048 (<= 256)");
    CSI_CHECK_ARG(!(drop_p > 0.f) || mask, "dropout backward needs the stored keep bits");
    const long long n = rows * C;
    const int grid = c2_grid(n >> 3);
    if (dtype == CSI_BF16)
        bn2d_bwd_apply_kernel<bf16><<<grid, C2_THREADS, 0, ST(stream)>>>(g, g_div, g_scale, (const bf16*)x, (const bf16*)zprev, mask, drop_p, n, C, mean,
                                                                          invstd, gamma, sums, 1.0 / (double)rows, (bf16*)gz, dgamma, dbeta);
    else
        bn2d_bwd_apply_kernel<float><<<grid, C2_THREADS, 0, ST(stream)>>>(g, g_div, g_scale, (const float*)x, (const float*)zprev, mask, drop_p, n, C, mean,
                                                                           invstd, gamma, sums, 1.0 / (double)rows, (float*)gz, dgamma, dbeta);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ final BatchNorm2d + mean over the image
// feat[b, c] = scale[c] * mean_p y[b, p, c] + shift[c]   (BatchNorm is affine per channel, so it commutes with the mean)
template <typename T>
__global__ void pool_bn_fwd_kernel(const T* __restrict__ y, int P, int C, const float* __restrict__ scale,
                                   const float* __restrict__ shift, float* __restrict__ feat, T* __restrict__ featd) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const T* src = y + (long long)b * P * C + c;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += ldv<T>(src + (long long)p * C);
        const float f = fmaf(s / (float)P, scale[c], shift[c]);
        feat[(long long)b * C + c] = f;
        stf<T>(featd + (long long)b * C + c, f);
    }
}

extern "C" int csi_pool_bn_fwd(const void* y, int dtype, int B, int P, int C, const float* scale, const float* shift, float* feat,
                               void* featd, void* stream) {
    CSI_CHECK_ARG(y && scale && shift && feat && featd && P >= 1 && C >= 1 && B >= 0, "bad argument");
    if (B == 0) return CSI_OK;
    if (dtype == CSI_BF16) pool_bn_fwd_kernel<bf16><<<B, 128, 0, ST(stream)>>>((const bf16*)y, P, C, scale, shift, feat, (bf16*)featd);
    else pool_bn_fwd_kernel<float><<<B, 128, 0, ST(stream)>>>((const float*)y, P, C, scale, shift, feat, (float*)featd);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ Conv2d weight re-layout
// reference layout w[n, c, kh, kw] (fp32 master) -> forward operand Wf[n, (kh*k + kw)*C + c] (ld = Kp, pad columns zero) and
// data-gradient operand Wb[(kh*k + kw)*C + c, n] (ld = Np, pad zero); and the inverse for the weight gradient.
template <typename T>
__global__ void conv2d_pack_kernel(const float* __restrict__ w, int N, int C, int k, T* __restrict__ wf, int Kp, T* __restrict__ wb, int Np) {
    const int K = k * k * C;
    const long long total = (long long)max(N, Np) * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / Kp), j = (int)(i - (long long)n * Kp);
        float v = 0.f;
        if (n < N && j < K) {
            const int c = j % C, t = j / C;                          // t = kh*k + kw
            v = w[((long long)n * C + c) * k * k + t];
        }
        if (n < N) stf<T>(wf + (long long)n * Kp + j, v);
        if (wb && n < Np) stf<T>(wb + (long long)j * Np + n, v);
    }
}
__global__ void conv2d_unpack_grad_kernel(const float* __restrict__ gs, int N, int C, int k, int Kp, float* __restrict__ gw) {
    const int K = k * k * C;
    const long long total = (long long)N * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / K), r = (int)(i - (long long)n * K);       // r = c*k*k + t in the reference layout
        const int c = r / (k * k), t = r - c * k * k;
        gw[i] += gs[(long long)n * Kp + t * C + c];
    }
}

extern "C" int csi_conv2d_pack(const float* w, int N, int C, int k, void* wf, int Kp, void* wb, int Np, int dtype, void* stream) {
    CSI_CHECK_ARG(w && wf && N >= 1 && C >= 1 && k >= 1 && Kp >= k * k * C && (!wb || Np >= N), "bad argument");
    const long long total = (long long)(wb ? (N > Np ? N : Np) : N) * Kp;
    const int grid = c2_grid(total);
    if (dtype == CSI_BF16) conv2d_pack_kernel<bf16><<<grid, C2_THREADS, 0, ST(stream)>>>(w, N, C, k, (bf16*)wf, Kp, (bf16*)wb, wb ? Np : N);
    else conv2d_pack_kernel<float><<<grid, C2_THREADS, 0, ST(stream)>>>(w, N, C, k, (float*)wf, Kp, (float*)wb, wb ? Np : N);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
extern "C" int csi_conv2d_unpack_grad(const float* gs, int N, int C, int k, int Kp, float* gw, void* stream) {
    CSI_CHECK_ARG(gs && gw && N >= 1 && C >= 1 && k >= 1 && Kp >= k * k * C, "bad argument");
    conv2d_unpack_grad_kernel<<<c2_grid((long long)N * k * k * C), C2_THREADS, 0, ST(stream)>>>(gs, N, C, k, Kp, gw);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ first BatchNorm2d (one channel)
// The image has ONE channel, so BatchNorm2d(1) in front of conv 0 is z0 = gamma0 * (W * xhat) + beta0 * sW + b with
// sW[n] = sum_k W[n, k]: its affine gradients follow from the column sums the backward pass has anyway,
//   dbeta0  = sum_n sW[n] * G[n],                      G[n]  = sum_m gz0[m, n]            (= the conv bias gradient)
//   dgamma0 = (sum_n GZ[n] - sum_n (b[n] + beta0 sW[n]) G[n]) / gamma0,   GZ[n] = sum_m gz0[m, n] * z0[m, n]
// instead of a 2.8 GB data-gradient GEMM + col2im over the input image.  (gamma0 == 0 makes z0 independent of the image;
// dgamma0 is then not recoverable from z0 and is reported as 0.)
template <typename T>
__global__ void bn0_grads_kernel(const double* __restrict__ sums, const T* __restrict__ wf, int ldw, int N, int K,
                                 const float* __restrict__ bias, const float* __restrict__ gamma0, const float* __restrict__ beta0,
                                 float* dgamma0, float* dbeta0, float* dbias) {
    __shared__ double acc[2];
    if (threadIdx.x == 0) acc[0] = acc[1] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int n = warp; n < N; n += nw) {
        float sw = 0.f;
        for (int j = lane; j < K; j += 32) sw += ldv<T>(wf + (long long)n * ldw + j);
        sw = warp_sum(sw);
        if (lane == 0) {
            const double G = sums[n], GZ = sums[N + n];
            dbias[n] += (float)G;
            atomicAdd(&acc[0], (double)sw * G);
            atomicAdd(&acc[1], GZ - ((double)bias[n] + (double)beta0[0] * sw) * G);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dbeta0[0] += (float)acc[0];
        const float g0 = gamma0[0];
        dgamma0[0] += g0 != 0.f ? (float)(acc[1] / (double)g0) : 0.f;
    }
}

extern "C" int csi_bn0_grads(const double* sums, const void* wf, int ldw, int dtype, int N, int K, const float* bias,
                             const float* gamma0, const float* beta0, float* dgamma0, float* dbeta0, float* dbias, void* stream) {
    CSI_CHECK_ARG(sums && wf && bias && gamma0 && beta0 && dgamma0 && dbeta0 && dbias && N >= 1 && K >= 1 && ldw >= K, "bad argument");
    if (dtype == CSI_BF16) bn0_grads_kernel<bf16><<<1, 256, 0, ST(stream)>>>(sums, (const bf16*)wf, ldw, N, K, bias, gamma0, beta0, dgamma0, dbeta0, dbias);
    else bn0_grads_kernel<float><<<1, 256, 0, ST(stream)>>>(sums, (const float*)wf, ldw, N, K, bias, gamma0, beta0, dgamma0, dbeta0, dbias);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// fp64 pools (statistics) are zeroed once per pass
__global__ void fill_f64_kernel(double* p, long long n, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
extern "C" int csi_fill_f64(double* p, long long n, double v, void* stream) {
    CSI_CHECK_ARG(p || n == 0, "null pointer");
    if (n == 0) return CSI_OK;
    fill_f64_kernel<<<c2_grid(n), C2_THREADS, 0, ST(stream)>>>(p, n, v);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

__global__ void copy_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
extern "C" int csi_copy_f32(float* dst, const float* src, long long n, void* stream) {
    CSI_CHECK_ARG((dst && src) || n == 0, "null pointer");
    if (n == 0) return CSI_OK;
    copy_f32_kernel<<<c2_grid(n), C2_THREADS, 0, ST(stream)>>>(dst, src, n);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

// ------------------------------------------------------------------------------------------------ batch gather (+ augmentation)
// The image path needs the dense, FRONT-padded batch (load_data.py:66-72) that the pooling kernel of the THAT path never
// materialises: out[b, t, :] = t >= T - lens[b] ? x[offs[b] + (t - (T - lens[b])) * F ...] : 0, with train.py:65-73
// (x + 0.1 N(0,1)) * U[0.9,1.1)_b * Bernoulli(0.96) applied on the way when `augment` (one Philox call per 4 elements:
// four 16-bit keep fields + two Box-Muller pairs from 16-bit uniforms).  offs == NULL: x is already dense [B, T, F].
#define SITE_IMG_AUG 9003u
#define SITE_IMG_SCALE 9004u
template <bool AUG>
__global__ void __launch_bounds__(C2_THREADS) gather_aug_kernel(const float* __restrict__ x, const long long* __restrict__ offs,
                                                                const int* __restrict__ lens, int T, int F, float* __restrict__ out,
                                                                long long total4, const unsigned long long* __restrict__ rng) {
    const long long per = (long long)T * F;                              // elements per sample (a multiple of 4)
    RngKey rk;
    if (AUG) rk = rng_load(rng);
    for (long long i4 = (long long)blockIdx.x * C2_THREADS + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * C2_THREADS) {
        const long long e = i4 * 4;
        const long long b = e / per, r = e - b * per;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (offs) {
            const long long padel = (long long)(T - lens[b]) * F;
            // a group of 4 never straddles the pad boundary when F % 4 == 0; otherwise fall back to scalars
            const float* src = x + offs[b] - padel;
            if ((padel & 3) == 0 && (offs[b] & 3) == 0) { if (r >= padel) v = *reinterpret_cast<const float4*>(src + r); }
            else {
                if (r >= padel) v.x = src[r];
                if (r + 1 >= padel) v.y = src[r + 1];
                if (r + 2 >= padel) v.z = src[r + 2];
                if (r + 3 >= padel) v.w = src[r + 3];
            }
        } else {
            v = *reinterpret_cast<const float4*>(x + e);
        }
        if (AUG) {
            const uint4 gs = rng_group(rk, SITE_IMG_SCALE, (unsigned long long)b);
            const float scale = (float)gs.x * (0.2f / 4294967296.0f) + 0.9f;                   // U[0.9, 1.1) per sample
            const uint4 g = rng_group(rk, SITE_IMG_AUG, (unsigned long long)i4);
            const uint32_t thr = 2621;                                                        // drop iff u16 < 0.04 * 65536
            const float u1 = ((float)(g.z & 0xFFFFu) + 1.0f) * (1.0f / 65536.0f), u2 = (float)(g.z >> 16) * (1.0f / 65536.0f);
            const float u3 = ((float)(g.w & 0xFFFFu) + 1.0f) * (1.0f / 65536.0f), u4 = (float)(g.w >> 16) * (1.0f / 65536.0f);
            const float ra = sqrtf(-2.0f * __logf(u1)) * 0.1f, rb = sqrtf(-2.0f * __logf(u3)) * 0.1f;
            float s0, c0, s1, c1;
            __sincosf(6.283185307179586f * u2, &s0, &c0);
            __sincosf(6.283185307179586f * u4, &s1, &c1);
            v.x = (g.x & 0xFFFFu) >= thr ? (v.x + ra * c0) * scale : 0.f;
            v.y = (g.x >> 16) >= thr ? (v.y + ra * s0) * scale : 0.f;
            v.z = (g.y & 0xFFFFu) >= thr ? (v.z + rb * c1) * scale : 0.f;
            v.w = (g.y >> 16) >= thr ? (v.w + rb * s1) * scale : 0.f;
        }
        *reinterpret_cast<float4*>(out + e) = v;
    }
}

extern "C" int csi_gather_aug(const float* x, const long long* offs, const int* lens, int B, int T, int F, float* out, int augment,
                              const unsigned long long* rng, void* stream) {
    CSI_CHECK_ARG(x && out && B >= 0 && T >= 1 && F >= 1, "bad argument");
    CSI_CHECK_ARG((offs == nullptr) == (lens == nullptr), "offs and lens come together");
    CSI_CHECK_ARG(((long long)T * F) % 4 == 0, "T*F must be a multiple of 4");
    CSI_CHECK_ARG(!augment || rng, "augmentation needs rng");
    const long long total4 = (long long)B * T * F / 4;
    if (total4 == 0) return CSI_OK;
    if (augment) gather_aug_kernel<true><<<c2_grid(total4), C2_THREADS, 0, ST(stream)>>>(x, offs, lens, T, F, out, total4, rng);
    else gather_aug_kernel<false><<<c2_grid(total4), C2_THREADS, 0, ST(stream)>>>(x, offs, lens, T, F, out, total4, rng);
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
