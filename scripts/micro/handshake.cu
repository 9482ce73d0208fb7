// Microbenchmark (diagnostic): cost of the pieces of the producer <-> MMA-issuer stage handshake.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diffusion-models_b200/csrc/ptx.cuh"
using namespace ddm;

// bit 0: wait on full   bit 1: tcgen05 fence after   bit 2: release with tcgen05.commit (else plain arrive if bit 3)
// bit 3: release with mbarrier.arrive   bit 4: issue 12 MMAs (N=64) per stage   bit 5: producer participates
template <int F>
__global__ void __launch_bounds__(128, 1) k(int stages, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8], empty[8];
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {           // producer
        if (F & 32) {
            int st = 0; uint32_t ph = 0;
            for (int i = 0; i < iters; ++i) {
                mbar_wait(&empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&full[st], 0);
                if (++st == stages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && (threadIdx.x & 31) == 0) {   // MMA issuer
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
        int st = 0; uint32_t ph = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (F & 1) mbar_wait(&full[st], ph);
            if (F & 2) tc_fence_after();
            if (F & 16) {
#pragma unroll
                for (int s = 0; s < 12; ++s) umma_bf16(tbase, a + 2u * (s & 3), b + 2u * (s & 3), idesc, 1);
            }
            if (F & 4) umma_commit(&empty[st]);
            if (F & 8) mbar_arrive(&empty[st]);
            if (++st == stages) { st = 0; ph ^= 1u; }
        }
        long long t1 = clock64();
        out[0] = (t1 - t0) / iters;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int F> void run(const char* what, long long* d) {
    cudaFuncSetAttribute(k<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<F><<<1, 128, 64 * 1024>>>(5, 2000, d);
    long long h; cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("MICRO %s: err %s\n", what, cudaGetErrorString(e)); exit(1); }
    printf("MICRO %-52s %5lld cycles/iteration\n", what, h);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0>("empty loop", d);
    run<2>("fence only", d);
    run<8>("arrive only", d);
    run<4>("commit only", d);
    run<16>("12 MMAs only", d);
    run<16 | 4>("12 MMAs + commit", d);
    run<32 | 1 | 8>("wait + arrive (full ring handshake)", d);
    run<32 | 1 | 2 | 8>("wait + fence + arrive", d);
    run<32 | 1 | 2 | 4>("wait + fence + commit", d);
    run<32 | 1 | 2 | 4 | 16>("wait + fence + 12 MMAs + commit", d);
    return 0;
}
