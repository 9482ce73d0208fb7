// Microbenchmark (diagnostic, not part of the library): issue rate of tcgen05.mma for small N, one vs several
// accumulators, and the latency of tcgen05.commit -> mbarrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diffusion-models_b200/csrc/ptx.cuh"
using namespace ddm;

template <int NACC>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
        uint32_t ph = 0;
        // warm
        umma_bf16(tbase, a, b, idesc, 0); umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
        long long t0 = clock64();
        const uint32_t d0 = tbase, d1 = tbase + (NACC > 1 ? N : 0), d2 = tbase + (NACC > 2 ? 2 * N : 0), d3 = tbase + (NACC > 2 ? 3 * N : 0);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int s = 0; s < 36; ++s) {
                const uint32_t d = (s % NACC) == 0 ? d0 : ((s % NACC) == 1 ? d1 : ((s % NACC) == 2 ? d2 : d3));
                umma_bf16(d, a + 2u * (s & 3), b + 2u * (s & 3), idesc, 1);
            }
        }
        umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
        long long t1 = clock64();
        out[0] = t1 - t0;
        // commit latency with nothing outstanding
        long long t2 = clock64();
        for (int it = 0; it < 64; ++it) { umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
        long long t3 = clock64();
        out[1] = (t3 - t2) / 64;
        // one MMA + commit + wait round trip
        long long t4 = clock64();
        for (int it = 0; it < 64; ++it) { umma_bf16(tbase, a, b, idesc, 1); umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
        long long t5 = clock64();
        out[2] = (t5 - t4) / 64;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int Ns[] = {64, 128, 256};
    for (int N : Ns) for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        const int iters = 64, ksteps = 36;
        if (nacc == 1) k<1><<<1, 128, 64 * 1024>>>(N, iters, d);
        else if (nacc == 2) k<2><<<1, 128, 64 * 1024>>>(N, iters, d);
        else k<4><<<1, 128, 64 * 1024>>>(N, iters, d);
        long long h[3]; cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        printf("MICRO N=%3d nacc=%d: %.1f cycles/MMA (ideal %d) | empty commit round trip %lld cyc | mma+commit+wait %lld cyc\n",
               N, nacc, double(h[0]) / (iters * ksteps), N / 2, h[1], h[2]);
    }
    return 0;
}
