// Microbenchmark (diagnostic): tcgen05.mma rate for N = 64/128/192/256 while other agents use shared memory:
//   mode 0 idle, mode 1 three warps stream ld/st.shared.v4, mode 2 one thread keeps bulk copies (global -> shared)
//   in flight (what the TMA producer does to the MMA's operand reads).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diffusion-models_b200/csrc/ptx.cuh"
using namespace ddm;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(160, 1) k(int N, int mode, int iters, const uint8_t* gsrc, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, cbar[4];
    __shared__ uint32_t tbase;
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1); fence_barrier_init(); done = 0; }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* scratch = smem + 49152;       // 96 KB of scratch after the operands
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int s = 0; s < 12; ++s) umma_bf16(tbase + (it & 1) * 256, a + 2u * (s & 3), b + 2u * (s & 3), idesc, 1);
        }
        umma_commit(&bar); mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        done = 1;
    } else if (warp >= 1 && warp <= 3 && mode == 1) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        uint32_t addr = smem_u32(scratch) + (warp - 1) * 16384 + lane * 16;
        long long n = 0;
        while (!done) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 v = lds_128u(addr + j * 512);
                acc.x ^= v.x; acc.y += v.y;
                sts_128u(addr + 8192 + j * 512, acc.x, acc.y, acc.z, acc.w);
            }
            ++n;
        }
        if (lane == 0 && blockIdx.x == 0) out[warp] = n * 8 * 2 * 512;    // bytes moved by this warp
    } else if (warp == 4 && lane == 0 && mode == 2) {
        long long n = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) { mbar_arrive_expect_tx(&cbar[i], 24576); bulk_g2s(scratch + i * 24576, gsrc + i * 24576, 24576, &cbar[i]); }
        while (!done) {
            for (int i = 0; i < 4; ++i) {
                mbar_wait(&cbar[i], ph[i]); ph[i] ^= 1;
                mbar_arrive_expect_tx(&cbar[i], 24576); bulk_g2s(scratch + i * 24576, gsrc + ((n * 4 + i) & 63) * 24576, 24576, &cbar[i]);
            }
            ++n;
        }
        for (int i = 0; i < 4; ++i) mbar_wait(&cbar[i], ph[i]);
        if (blockIdx.x == 0) out[1] = n * 4 * 24576;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    const int iters = argc > 2 ? atoi(argv[2]) : 400;
    long long* d; cudaMalloc(&d, 64);
    uint8_t* g; cudaMalloc(&g, 64 * 24576 + 4096); cudaMemset(g, 0, 64 * 24576);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {64, 128, 192, 256}) for (int mode : {0, 1, 2}) {
        cudaMemset(d, 0, 64);
        k<<<grid, 160, 160 * 1024>>>(N, mode, iters, g, d);
        long long h[4]; cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("MICRO err %s\n", cudaGetErrorString(e)); return 1; }
        const double other = double(h[1] + h[2] + h[3]);
        printf("MICRO N=%3d mode=%d: %.1f cycles/MMA (math %d, operand bytes %d) | other smem traffic %.1f B/clk\n", N, mode,
               double(h[0]) / (iters * 12), N / 2, 4096 + N * 32, other / double(h[0]));
    }
    return 0;
}
