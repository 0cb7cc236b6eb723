// Per-head attention core of nn.MultiheadAttention (that.py:149): softmax(q k^T / sqrt(hd)) v, forward and
// backward, one CTA per (sample, head) with the whole head's K/V (and Q/dO for backward) resident in shared
// memory; scores never touch HBM (the reference materialises [B,10,L,L] probabilities).  fp32 math.
#include "common.cuh"

#define ST(s) ((cudaStream_t)(s))
#define ATT_WARPS 8
#define ATT_JMAX 20           // keys per lane: L <= 640

template <typename T>
__device__ __forceinline__ void load_head(const T* __restrict__ src, int ld, int col0, int L, int hd, int hs,
                                          float* __restrict__ dst) {
    for (int i = threadIdx.x; i < L * hd; i += blockDim.x) {
        const int l = i / hd, e = i % hd;
        dst[l * hs + e] = ldv<T>(src + (size_t)l * ld + col0 + e);
    }
}

// out[e] (e = lane % W (+32)) = sum_j pw[j] * M[j][e]; lanes are split in 32/W groups over j
template <int DUMMY = 0>
__device__ __forceinline__ void row_times_matrix(const float* __restrict__ pw, const float* __restrict__ Mx, int L,
                                                 int hd, int hs, int W, int lane, float& o0, float& o1) {
    const int e0 = lane % W, grp = lane / W, ng = 32 / W;
    float a0 = 0.f, a1 = 0.f;
    const bool v0 = e0 < hd, v1 = (e0 + 32) < hd;
    for (int j = grp; j < L; j += ng) {
        const float pj = pw[j];
        if (v0) a0 = fmaf(pj, Mx[j * hs + e0], a0);
        if (v1) a1 = fmaf(pj, Mx[j * hs + e0 + 32], a1);
    }
    for (int o = W; o < 32; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    o0 = a0; o1 = a1;
}

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_fwd_kernel(const T* __restrict__ qkv, int ld3, T* __restrict__ o,
                                                                  int ldo, float* __restrict__ lse, int L, int d, int H,
                                                                  int halo, int hs, int W, int hp) {
    extern __shared__ float sm[];
    const int hd = d / H, b = blockIdx.x / H, h = blockIdx.x % H, Lp = L + 2 * halo;
    float* Ks = sm;
    float* Vs = Ks + L * hs;
    float* Ps = Vs + L * hs;                 // [ATT_WARPS][L]
    float* Qs = Ps + ATT_WARPS * L;          // [ATT_WARPS][hs]
    const size_t row0 = (size_t)b * Lp + halo;
    const T* base = qkv + row0 * ld3;
    load_head<T>(base, ld3, H * hp + h * hp, L, hd, hs, Ks);
    load_head<T>(base, ld3, 2 * H * hp + h * hp, L, hd, hs, Vs);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float sc = rsqrtf((float)hd);
    float* pw = Ps + wid * L;
    float* qw = Qs + wid * hs;
    for (int i = wid; i < L; i += ATT_WARPS) {
        for (int e = lane; e < hd; e += 32) qw[e] = ldv<T>(base + (size_t)i * ld3 + h * hp + e) * sc;
        __syncwarp();
        float s[ATT_JMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < ATT_JMAX; ++jj) {
            const int j = lane + 32 * jj;
            float a = -INFINITY;
            if (j < L) {
                a = 0.f;
                for (int e = 0; e < hd; ++e) a = fmaf(qw[e], Ks[j * hs + e], a);
            }
            s[jj] = a;
            mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < ATT_JMAX; ++jj) {
            const int j = lane + 32 * jj;
            if (j < L) { const float p = __expf(s[jj] - mx); pw[j] = p; sum += p; }
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0, o1;
        row_times_matrix(pw, Vs, L, hd, hs, W, lane, o0, o1);
        const float inv = 1.f / sum;
        const int e0 = lane % W;
        if (lane < W) {
            T* orow = o + (row0 + i) * ldo + h * hp;
            if (e0 < hd) stf<T>(orow + e0, o0 * inv);
            if (e0 + 32 < hd) stf<T>(orow + e0 + 32, o1 * inv);
        }
        if (lane == 0) lse[((size_t)b * H + h) * L + i] = mx + __logf(sum);
        __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_bwd_kernel(const T* __restrict__ qkv, int ld3,
                                                                  const T* __restrict__ o, int ldo,
                                                                  const T* __restrict__ dout, int lddo,
                                                                  T* __restrict__ dqkv, int lddqkv,
                                                                  const float* __restrict__ lse, int L, int d, int H,
                                                                  int halo, int hs, int W, int hp) {
    extern __shared__ float sm[];
    const int hd = d / H, b = blockIdx.x / H, h = blockIdx.x % H, Lp = L + 2 * halo;
    float* Qs = sm;
    float* Ks = Qs + L * hs;
    float* Vs = Ks + L * hs;
    float* Gs = Vs + L * hs;                 // dO
    float* Ls = Gs + L * hs;                 // lse [L]
    float* Ds = Ls + L;                      // rowsum(dO*O) [L]
    float* Ps = Ds + L;                      // [ATT_WARPS][L]
    float* Ss = Ps + ATT_WARPS * L;          // [ATT_WARPS][L]
    const size_t row0 = (size_t)b * Lp + halo;
    const T* base = qkv + row0 * ld3;
    load_head<T>(base, ld3, h * hp, L, hd, hs, Qs);
    load_head<T>(base, ld3, H * hp + h * hp, L, hd, hs, Ks);
    load_head<T>(base, ld3, 2 * H * hp + h * hp, L, hd, hs, Vs);
    load_head<T>(dout + row0 * lddo, lddo, h * hp, L, hd, hs, Gs);
    for (int i = threadIdx.x; i < L; i += blockDim.x) Ls[i] = lse[((size_t)b * H + h) * L + i];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        float a = 0.f;
        for (int e = 0; e < hd; ++e) a = fmaf(Gs[i * hs + e], ldv<T>(o + (row0 + i) * ldo + h * hp + e), a);
        Ds[i] = a;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float sc = rsqrtf((float)hd);
    float* pw = Ps + wid * L;
    float* sw = Ss + wid * L;
    const int e0 = lane % W;
    // phase A: dQ_i = sc * sum_j dS_ij K_j, one query row per warp iteration
    for (int i = wid; i < L; i += ATT_WARPS) {
        const float li = Ls[i], di = Ds[i];
        for (int j = lane; j < L; j += 32) {
            float s = 0.f, dp = 0.f;
            for (int e = 0; e < hd; ++e) {
                s = fmaf(Qs[i * hs + e], Ks[j * hs + e], s);
                dp = fmaf(Gs[i * hs + e], Vs[j * hs + e], dp);
            }
            const float p = __expf(s * sc - li);
            sw[j] = p * (dp - di) * sc;
        }
        __syncwarp();
        float o0, o1;
        row_times_matrix(sw, Ks, L, hd, hs, W, lane, o0, o1);
        if (lane < W) {
            T* dq = dqkv + (row0 + i) * lddqkv + h * hp;
            if (e0 < hd) stf<T>(dq + e0, o0);
            if (e0 + 32 < hd) stf<T>(dq + e0 + 32, o1);
        }
        __syncwarp();
    }
    // phase B: dK_j = sc * sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i, one key row per warp iteration
    for (int j = wid; j < L; j += ATT_WARPS) {
        for (int i = lane; i < L; i += 32) {
            float s = 0.f, dp = 0.f;
            for (int e = 0; e < hd; ++e) {
                s = fmaf(Qs[i * hs + e], Ks[j * hs + e], s);
                dp = fmaf(Gs[i * hs + e], Vs[j * hs + e], dp);
            }
            const float p = __expf(s * sc - Ls[i]);
            pw[i] = p;
            sw[i] = p * (dp - Ds[i]) * sc;
        }
        __syncwarp();
        float k0, k1, v0, v1;
        row_times_matrix(sw, Qs, L, hd, hs, W, lane, k0, k1);
        row_times_matrix(pw, Gs, L, hd, hs, W, lane, v0, v1);
        if (lane < W) {
            T* dk = dqkv + (row0 + j) * lddqkv + H * hp + h * hp;
            T* dv = dqkv + (row0 + j) * lddqkv + 2 * H * hp + h * hp;
            if (e0 < hd) { stf<T>(dk + e0, k0); stf<T>(dv + e0, v0); }
            if (e0 + 32 < hd) { stf<T>(dk + e0 + 32, k1); stf<T>(dv + e0 + 32, v1); }
        }
        __syncwarp();
    }
}

static int head_stride(int hd) { return hd | 1; }                  // odd stride: conflict-free row-strided reads
static int lane_width(int hd) { int w = 1; while (w < hd && w < 32) w <<= 1; return w; }

extern "C" int csi_attn_fwd_simt(const void* qkv, int ld3, void* o, int ldo, int dtype, float* lse, int B, int L, int d,
                                 int H, int hp, int halo, void* stream) {
    CSI_CHECK_ARG(qkv && o && lse, "null pointer");
    CSI_CHECK_ARG(H > 0 && d % H == 0 && d / H <= 64 && L <= 32 * ATT_JMAX, "head_dim <= 64 and L <= 640");
    if (B == 0) return CSI_OK;
    const int hd = d / H, hs = head_stride(hd), W = lane_width(hd);
    const size_t smem = ((size_t)2 * L * hs + (size_t)ATT_WARPS * L + ATT_WARPS * hs) * sizeof(float);
    CSI_CHECK_ARG(smem <= 227 * 1024, "head does not fit in shared memory");
    if (dtype == CSI_BF16) {
        CSI_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_fwd_kernel<bf16><<<B * H, ATT_WARPS * 32, smem, ST(stream)>>>((const bf16*)qkv, ld3, (bf16*)o, ldo, lse, L, d, H, halo, hs, W, hp);
    } else {
        CSI_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_fwd_kernel<float><<<B * H, ATT_WARPS * 32, smem, ST(stream)>>>((const float*)qkv, ld3, (float*)o, ldo, lse, L, d, H, halo, hs, W, hp);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}

extern "C" int csi_attn_bwd_simt(const void* qkv, int ld3, const void* o, int ldo, const void* dout, int lddo,
                                 void* dqkv, int lddqkv, int dtype, const float* lse, int B, int L, int d, int H,
                                 int hp, int halo, void* stream) {
    CSI_CHECK_ARG(qkv && o && dout && dqkv && lse, "null pointer");
    CSI_CHECK_ARG(H > 0 && d % H == 0 && d / H <= 64, "head_dim <= 64");
    if (B == 0) return CSI_OK;
    const int hd = d / H, hs = head_stride(hd), W = lane_width(hd);
    const size_t smem = ((size_t)4 * L * hs + 2 * (size_t)L + 2 * (size_t)ATT_WARPS * L) * sizeof(float);
    CSI_CHECK_ARG(smem <= 227 * 1024, "head does not fit in shared memory");
    if (dtype == CSI_BF16) {
        CSI_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_bwd_kernel<bf16><<<B * H, ATT_WARPS * 32, smem, ST(stream)>>>((const bf16*)qkv, ld3, (const bf16*)o, ldo, (const bf16*)dout,
                                                                         lddo, (bf16*)dqkv, lddqkv, lse, L, d, H, halo, hs, W, hp);
    } else {
        CSI_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_bwd_kernel<float><<<B * H, ATT_WARPS * 32, smem, ST(stream)>>>((const float*)qkv, ld3, (const float*)o, ldo,
                                                                          (const float*)dout, lddo, (float*)dqkv, lddqkv, lse, L, d, H,
                                                                          halo, hs, W, hp);
    }
    CSI_LAUNCH_CHECK();
    return CSI_OK;
}
