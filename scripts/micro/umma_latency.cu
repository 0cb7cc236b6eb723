// Microbenchmark: cost of a tcgen05.mma (M=128, K=16, bf16) as a function of N and of how many independent TMEM
// accumulators the instruction stream rotates over.  One CTA; operands are whatever lies in shared memory (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_latency umma_latency.cu && ./umma_latency
#include "../../multi_modal_csi_b200/csrc/tc_common.cuh"
#include <cstdio>
void csi_set_error(const char*, ...) {}

__device__ __forceinline__ void umma_ts_(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MODE 0: A, B K-major from smem; 1: A from TMEM, B MN-major; 2: A and B MN-major; 3: A K-major, B MN-major
template <int NACC, int MODE>
__global__ void __launch_bounds__(128, 1) k_umma(int N, int nmma, int a_stride, int b_stride, long long* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    fence_proxy_async();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(128, N) | (MODE == 2 ? (1u << 15) : 0u) | (MODE >= 1 ? (1u << 16) : 0u);
        const uint64_t a0 = MODE == 2 ? mn_desc(smem_u32(smem), 16384) : make_kmajor_desc(smem_u32(smem));
        const uint64_t b0 = MODE >= 1 ? mn_desc(smem_u32(smem) + 65536, 16384) : make_kmajor_desc(smem_u32(smem) + 65536);
        constexpr int pitch = 512 / NACC;
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = clock64();
            uint64_t a = a0, b = b0;
            for (int i = 0; i < nmma; i += 4 * NACC) {           // straight-line groups: no division, constants only
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int acc = 0; acc < NACC; ++acc)
                        if (MODE == 1) umma_ts_(tbase + acc * pitch, tbase + 256 + u * 8, b + (uint64_t)(u * b_stride), idesc, 1u);
                        else umma_bf16(tbase + acc * pitch, a + (uint64_t)(u * a_stride), b + (uint64_t)(u * b_stride), idesc, 1u);
            }
            long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int MODE> void run(long long* d, const char* name) {
    cudaFuncSetAttribute(k_umma<1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_umma<2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int nmma = 128;
    for (int N : {16, 32, 64, 96, 128, 160, 256})
        for (int nacc : {1, 2}) {
            if (N * nacc > 256) continue;
            if (nacc == 1) k_umma<1, MODE><<<1, 128, 200 * 1024>>>(N, nmma, 2, 128, d);
            else k_umma<2, MODE><<<1, 128, 200 * 1024>>>(N, nmma, 2, 128, d);
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("%-28s N=%3d nacc=%d  issue %6.1f  total %6.1f clk/UMMA %s\n", name, N, nacc, (double)h[0] / nmma, (double)h[1] / nmma, e ? cudaGetErrorString(e) : "");
        }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    run<0>(d, "SS  A K-major, B K-major");
    run<3>(d, "SS  A K-major, B MN-major");
    run<1>(d, "TS  A tmem,    B MN-major");
    run<2>(d, "SS  A MN-major, B MN-major");
    return 0;
}
