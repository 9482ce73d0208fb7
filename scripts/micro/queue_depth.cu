// Microbenchmark (diagnostic): how many tcgen05.mma can be issued before the issuing thread blocks (queue depth),
// and the cost of mbarrier.try_wait on an already-completed phase.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diffusion-models_b200/csrc/ptx.cuh"
using namespace ddm;

template <int K>
__global__ void __launch_bounds__(128, 1) k(int N, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, done;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
        uint32_t ph = 0;
        umma_bf16(tbase, a, b, idesc, 0); umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;   // warm
        long long issue = 0, total = 0;
        for (int rep = 0; rep < 16; ++rep) {
            long long t0 = clock64();
#pragma unroll
            for (int s = 0; s < K; ++s) umma_bf16(tbase, a + 2u * (s & 3), b + 2u * (s & 3), idesc, 1);
            long long t1 = clock64();
            umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
            long long t2 = clock64();
            issue += t1 - t0; total += t2 - t0;
        }
        out[0] = issue / 16; out[1] = total / 16;
        // try_wait on a long-completed phase (parity of the previous phase)
        long long t3 = clock64();
        for (int i = 0; i < 256; ++i) mbar_wait(&bar, ph ^ 1);
        long long t4 = clock64();
        out[2] = (t4 - t3) / 256;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int K> void run(int N, long long* d) {
    cudaFuncSetAttribute(k<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<K><<<1, 128, 64 * 1024>>>(N, d);
    long long h[3]; cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("MICRO err %s\n", cudaGetErrorString(e)); exit(1); }
    printf("MICRO N=%3d K=%2d MMAs: issue %5lld cyc, issue+commit+wait %5lld cyc (pipe ideal %4d) | completed try_wait %lld cyc\n",
           N, K, h[0], h[1], K * (N == 64 ? 48 : N / 2), h[2]);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int N : {64, 256}) { run<1>(N, d); run<2>(N, d); run<4>(N, d); run<8>(N, d); run<16>(N, d); run<32>(N, d); }
    return 0;
}
