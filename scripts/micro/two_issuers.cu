// Microbenchmark (diagnostic): aggregate tcgen05.mma rate with one vs two concurrently issuing threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diffusion-models_b200/csrc/ptx.cuh"
using namespace ddm;

__global__ void __launch_bounds__(128, 1) k(int N, int nissuers, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0 && warp < nissuers) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
        const uint32_t d = tbase + warp * N;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int s = 0; s < 12; ++s) umma_bf16(d, a + 2u * (s & 3), b + 2u * (s & 3), idesc, 1);
        }
        umma_commit(&bar[warp]); mbar_wait(&bar[warp], 0);
        long long t1 = clock64();
        out[warp] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int N : {64, 128, 256}) for (int ni : {1, 2}) {
        const int iters = 200;
        k<<<1, 128, 64 * 1024>>>(N, ni, iters, d);
        long long h[2] = {0, 0}; cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("MICRO err %s\n", cudaGetErrorString(e)); return 1; }
        const double total_mmas = double(ni) * iters * 12;
        const long long t = h[0] > h[1] ? h[0] : h[1];
        printf("MICRO N=%3d issuers=%d: %.1f cycles per MMA aggregate (pipe ideal %d)\n", N, ni, t / total_mmas, N == 64 ? 48 : N / 2);
    }
    return 0;
}
