// Stem convolution (init_conv 7x7, pad 3; denoising_diffusion.py:262,356) as an implicit GEMM on mma.sync:
//   M = 128 pixels (8 rows x 16 columns of one image) per CTA, N = C_out, K = ks*ks*C_in padded to 16.
// C_in is 3..8, so a tcgen05/TMA formulation has nothing to tile along channels; instead the fp32 NCHW input patch is
// staged once in shared memory as bf16, the A fragments are gathered from it through a k -> patch-offset table, and
// the weights sit in shared memory as the B operand.  Output: bf16 channels-last, written as full 2*C_out-byte rows.
#include "kernels.cuh"

#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {
namespace {

constexpr int TH = 8, TW = 16;       // pixel tile
constexpr int WPAD = 8;              // bf16 padding of weight / output rows (conflict-free ldmatrix)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>   // NT = C_out / 8 n-tiles (even)
__global__ void __launch_bounds__(128)
stem_tc_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1, const float* __restrict__ in2,
               int c2, const float* __restrict__ weight, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
               int H, int W, int ks, int k_pad, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int COUT = NT * 8;
    constexpr int WP = COUT + WPAD;
    extern __shared__ __align__(16) uint8_t stem_smem[];
    const int Cin = c0 + c1 + c2;
    const int K = ks * ks * Cin;
    const int pad = ks / 2;
    const int PH = TH + ks - 1, PW = TW + ks - 1;
    __nv_bfloat16* w_s = reinterpret_cast<__nv_bfloat16*>(stem_smem);                 // [k_pad][WP]
    int* koff = reinterpret_cast<int*>(w_s + k_pad * WP);                              // [k_pad]
    int* pidx = koff + k_pad;                                                          // [Cin*PH*PW] packed (src, c, py, px)
    const int n_patch = Cin * PH * PW;
    __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(pidx + ((n_patch + 3) & ~3));  // [Cin][PH][PW]
    __nv_bfloat16* o_s = patch + ((Cin * PH * PW + 7) & ~7);                           // [128][WP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < k_pad * COUT; i += 128) {
        const int k = i / COUT, n = i - k * COUT;
        w_s[k * WP + n] = __float2bfloat16_rn(k < K ? __ldg(weight + k * COUT + n) : 0.0f);
    }
    for (int k = tid; k < k_pad; k += 128) {
        int off = 0;
        if (k < K) {                                    // weight rows are tap-major: k = (ky*ks + kx)*Cin + ci
            const int tap = k / Cin, ci = k - tap * Cin;
            const int ky = tap / ks, kx = tap - ky * ks;
            off = (ci * PH + ky) * PW + kx;
        }
        koff[k] = off;
    }
    // element i of the patch -> (source, channel in source, patch row, patch column), decoded once per CTA
    for (int i = tid; i < n_patch; i += 128) {
        const int ci = i / (PH * PW);
        const int rem = i - ci * PH * PW;
        const int py = rem / PW, px = rem - py * PW;
        int src = 0, c = ci;
        if (ci >= c0 + c1) { src = 2; c = ci - c0 - c1; }
        else if (ci >= c0) { src = 1; c = ci - c0; }
        pidx[i] = (src << 24) | (c << 16) | (py << 8) | px;
    }
    __syncthreads();
    // persistent CTA: weights and tables are staged once, then the CTA walks its tiles.  The fp32 patch of the NEXT
    // tile is fetched into registers while the current tile is multiplied (the kernel was bound by that latency).
    constexpr int PRE = 20;                        // 128 * 20 >= 8 channels x 14 x 22
    const int iters = (n_patch + 127) / 128;       // host guarantees iters <= PRE
    float pre[PRE];
    auto fetch_patch = [&](int tile) {
        const int b = tile / (tiles_x * tiles_y);
        const int trem = tile - b * (tiles_x * tiles_y);
        const int y0 = (trem / tiles_x) * TH, x0 = (trem % tiles_x) * TW;
#pragma unroll
        for (int it = 0; it < PRE; ++it) {
            const int i = tid + it * 128;
            float v = 0.0f;
            if (it < iters && i < n_patch) {
                const int e = pidx[i];
                const int y = y0 + ((e >> 8) & 0xFF) - pad, x = x0 + (e & 0xFF) - pad;
                if (y >= 0 && y < H && x >= 0 && x < W) {
                    const int sel = e >> 24, c = (e >> 16) & 0xFF;
                    const float* src = sel == 0 ? in0 : (sel == 1 ? in1 : in2);
                    const int cn = sel == 0 ? c0 : (sel == 1 ? c1 : c2);
                    v = __ldg(src + ((static_cast<long long>(b) * cn + c) * H + y) * W + x);
                }
            }
            pre[it] = v;
        }
    };
    if (static_cast<int>(blockIdx.x) < total_tiles) fetch_patch(blockIdx.x);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = tile / (tiles_x * tiles_y);
    const int trem = tile - b * (tiles_x * tiles_y);
    const int y0 = (trem / tiles_x) * TH, x0 = (trem % tiles_x) * TW;
#pragma unroll
    for (int it = 0; it < PRE; ++it) {
        const int i = tid + it * 128;
        if (it < iters && i < n_patch) patch[i] = __float2bfloat16_rn(pre[it]);
    }
    if (tile + static_cast<int>(gridDim.x) < total_tiles) fetch_patch(tile + gridDim.x);
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
    // warp w owns tile rows 2w, 2w+1 (one m-tile = one row of 16 pixels); fragment rows g / g+8 = pixel columns
    const unsigned short* pu = reinterpret_cast<const unsigned short*>(patch);
    for (int k0 = 0; k0 < k_pad; k0 += 16) {
        const int2 ka = *reinterpret_cast<const int2*>(koff + k0 + 2 * t);        // k = k0+2t, k0+2t+1
        const int2 kb = *reinterpret_cast<const int2*>(koff + k0 + 8 + 2 * t);    // k = k0+8+2t, +1
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int base = (warp * 2 + mt) * PW + g;
            a[mt][0] = pu[base + ka.x] | (static_cast<uint32_t>(pu[base + ka.y]) << 16);
            a[mt][1] = pu[base + 8 + ka.x] | (static_cast<uint32_t>(pu[base + 8 + ka.y]) << 16);
            a[mt][2] = pu[base + kb.x] | (static_cast<uint32_t>(pu[base + kb.y]) << 16);
            a[mt][3] = pu[base + 8 + kb.x] | (static_cast<uint32_t>(pu[base + 8 + kb.y]) << 16);
        }
        const int i = lane >> 3, j = lane & 7;
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, w_s + (k0 + j + 8 * (i & 1)) * WP + np * 16 + 8 * (i >> 1));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_bf16(acc[mt][2 * np], a[mt], bf[0], bf[1]);
                mma_bf16(acc[mt][2 * np + 1], a[mt], bf[2], bf[3]);
            }
        }
    }
    // bias, bf16, stage as [pixel][C_out] so that each thread then writes one full output row
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int p0 = (warp * 2 + mt) * TW + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int n = nt * 8 + 2 * t;
            const float b0 = __ldg(bias + n), b1 = __ldg(bias + n + 1);
            *reinterpret_cast<uint32_t*>(o_s + p0 * WP + n) = pack_bf16x2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
            *reinterpret_cast<uint32_t*>(o_s + (p0 + 8) * WP + n) = pack_bf16x2(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
        }
    }
    __syncthreads();
    const int py = tid / TW, px = tid - py * TW;
    const int y = y0 + py, x = x0 + px;
    if (y < H && x < W) {
        const uint4* src = reinterpret_cast<const uint4*>(o_s + tid * WP);
        uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * H + y) * W + x) * COUT);
#pragma unroll
        for (int c = 0; c < COUT / 8; ++c) dst[c] = src[c];
    }
    __syncthreads();        // patch / output staging are rewritten by the next tile
    }
}

int stem_tc_smem(int Cin, int Cout, int ks, int k_pad) {
    const int PH = TH + ks - 1, PW = TW + ks - 1;
    return k_pad * (Cout + WPAD) * 2 + k_pad * 4 + ((Cin * PH * PW + 3) & ~3) * 4 + ((Cin * PH * PW + 7) & ~7) * 2 + 128 * (Cout + WPAD) * 2;
}

}  // namespace

bool stem_tc_supported(int Cin, int Cout, int ks) {
    const int k_pad = (ks * ks * Cin + 15) / 16 * 16;
    const int n_patch = Cin * (TH + ks - 1) * (TW + ks - 1);
    return (Cout == 32 || Cout == 64 || Cout == 128) && n_patch <= 128 * 20 && Cin < 256 && (TH + ks - 1) < 256 && (TW + ks - 1) < 256 &&
           stem_tc_smem(Cin, Cout, ks, k_pad) <= 200 * 1024;
}

int stem_tc_prepare_attributes() {
    int r = static_cast<int>(cudaFuncSetAttribute(stem_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(stem_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(stem_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return r;
}

void launch_stem_tc(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                    void* out, int B, int H, int W, int Cout, int ks, int num_sms, cudaStream_t s) {
    const int Cin = c0 + c1 + c2;
    const int k_pad = (ks * ks * Cin + 15) / 16 * 16;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const int total = tiles_x * tiles_y * B;
    const int smem = stem_tc_smem(Cin, Cout, ks, k_pad);
    int per_sm = (220 * 1024) / (smem + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = total < num_sms * per_sm ? total : num_sms * per_sm;
    auto* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (Cout == 32) stem_tc_kernel<4><<<grid, 128, smem, s>>>(in0, c0, in1, c1, in2, c2, w, b, o, H, W, ks, k_pad, tiles_x, tiles_y, total);
    else if (Cout == 64) stem_tc_kernel<8><<<grid, 128, smem, s>>>(in0, c0, in1, c1, in2, c2, w, b, o, H, W, ks, k_pad, tiles_x, tiles_y, total);
    else stem_tc_kernel<16><<<grid, 128, smem, s>>>(in0, c0, in1, c1, in2, c2, w, b, o, H, W, ks, k_pad, tiles_x, tiles_y, total);
}

}  // namespace ddm
