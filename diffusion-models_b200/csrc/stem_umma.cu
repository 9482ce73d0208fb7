// Stem convolution (init_conv 7x7, pad 3; denoising_diffusion.py:262,356) on tcgen05 for few input channels (C_in <= 4).
//
// The 3-channel input offers nothing to tile along channels, so the A operand (128 pixels x K) is an explicit im2col tile that the
// CTA's threads assemble in shared memory from an fp32 patch of the NCHW input -- in the 128-byte-swizzled K-major layout the UMMA
// descriptor reads.  K is ordered (channel, ky, kx) with kx padded from 7 to 8: one 16-byte unit of an A row is the 7
// horizontally adjacent patch values of one (channel, ky) row plus a zero, i.e. 7 conflict-free 4-byte shared loads (lanes =
// adjacent pixels), 4 packs and one 16-byte store.  C_in = 3: K = 3 x 7 x 8 = 168 -> 192 (three swizzle atoms), 12 MMAs of
// M = 128, N = C_out per 128-pixel tile.  The weights are repacked the same way once per CTA and stay in shared memory.
//
// Two persistent CTAs per SM, 256 threads each (128 registers: the prefetched patch must not spill -- a spill store waits for its load), no warp specialisation and nothing overlapped inside a CTA: a tile is assembled,
// thread 0 issues its MMAs (0.6 k cycles) and all threads run the epilogue (bias, bf16, staged so that the global stores are full
// 128-byte rows); the fp32 patch of the CTA's next tile is in flight in registers meanwhile, and the SM's other CTA fills the
// bubbles.  (A first version -- one 256-thread CTA per SM, A and the accumulator double buffered -- was latency-bound at two
// warps per scheduler: 196 us at B = 1024, 32 x 32.)
// (The mma.sync stem in stem_tc.cu gathers fragments through 2-byte shared loads: 141 us at B = 1024, 32 x 32; it remains the
// path for C_in > 4 -- self-conditioning, image-conditional nets -- and other kernel sizes.)
#include "kernels.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {
namespace {

constexpr int kThreads = 256;
constexpr int kKs = 7, kPad = 3;
constexpr int kPre = 6;                 // patch values prefetched per thread: 256 x 6 >= 4 channels x 10 x 38

struct StemBars {
    uint64_t mma_done;
    uint32_t tmem_base;
};

struct StemPlan {
    int atoms, pitch, w_off, a_off, patch_off, joff_off, stg_off, bias_off, bars_off, total;
};
__host__ __device__ inline StemPlan stem_plan(int Cin, int Cout, int TW) {
    const int TH = 128 / TW;
    StemPlan s;
    s.atoms = (Cin * kKs + 7) / 8;                       // 64-element K atoms (8 units of (channel, ky))
    s.pitch = TW + 8;                                    // bf16 elements per patch row (TW + 6 used; rows stay 4-byte aligned)
    s.w_off = 0;
    s.a_off = s.w_off + s.atoms * Cout * 128;
    s.patch_off = s.a_off + s.atoms * 128 * 128;
    s.joff_off = s.patch_off + ((Cin * (TH + 6) * s.pitch * 2 + 15) & ~15);
    s.stg_off = (s.joff_off + 32 * 4 + 1023) & ~1023;
    s.bias_off = s.stg_off + 128 * Cout * 2;
    s.bars_off = s.bias_off + Cout * 4;
    s.total = s.bars_off + static_cast<int>(sizeof(StemBars));
    return s;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(addr), "h"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
stem_umma_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1, const float* __restrict__ in2, int c2,
                 const float* __restrict__ weight, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                 int H, int W, int Cout, int TW, int tw_shift, int tiles_x, int tiles_y, int total_tiles, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t stem_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stem_smem_raw) + 1023) & ~uintptr_t(1023));
    const int Cin = c0 + c1 + c2;
    const int TH = 128 >> tw_shift;
    const int PH = TH + 6, PW = TW + 6;
    const StemPlan pl = stem_plan(Cin, Cout, TW);
    const uint32_t sb = smem_u32(smem);
    float* bias_s = reinterpret_cast<float*>(smem + pl.bias_off);
    int* joff_s = reinterpret_cast<int*>(smem + pl.joff_off);           // (both written once, then read with shared-space loads)
    StemBars* bars = reinterpret_cast<StemBars*>(smem + pl.bars_off);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nJ = Cin * kKs;                            // 16-byte units of an A row that carry data
    const int n_patch = Cin * PH * PW;

    griddep_launch();
    if (tid == 0) {
        mbar_init(&bars->mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&bars->tmem_base, static_cast<uint32_t>(tmem_cols));
        tmem_relinquish();
    }
    // weights: global fp32 [k = (ky*7 + kx)*Cin + ci][Cout]  ->  bf16 [atom][n][unit j%8 ^ (n&7)][kx], j = ci*7 + ky
    for (int i = tid; i < pl.atoms * 8 * Cout; i += kThreads) {
        const int n = i % Cout, j = i / Cout;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (j < nJ) {
            const int ci = j / kKs, ky = j - ci * kKs;
            float f[8];
#pragma unroll
            for (int kx = 0; kx < 8; ++kx)
                f[kx] = kx < kKs ? __ldg(weight + static_cast<long long>((ky * kKs + kx) * Cin + ci) * Cout + n) : 0.0f;
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack2(f[2 * e], f[2 * e + 1]);
        }
        sts_128u(sb + pl.w_off + (j >> 3) * (Cout * 128) + n * 128 + (((j & 7) ^ (n & 7)) << 4), w[0], w[1], w[2], w[3]);
    }
    // the units of the last atom beyond nJ are never written by the tile loop: zero the A buffer once
    for (int i = tid; i < pl.atoms * 128 * 8; i += kThreads) sts_128u(sb + pl.a_off + i * 16, 0u, 0u, 0u, 0u);
    for (int i = tid; i < Cout; i += kThreads) bias_s[i] = __ldg(bias + i);
    if (tid < 32) {           // unit j = (ci, ky) -> byte offset of its patch row
        const int ci = tid / kKs, ky = tid - ci * kKs;
        joff_s[tid] = (ci * PH + ky) * pl.pitch * 2;
    }
    // this thread's patch elements, decoded once: offset of the element inside its source for image 0 / tile origin (0, 0)
    // (possibly negative; dereferenced only when the coordinates are inside), and (shared-memory byte offset, source, py, px)
    int eoff[kPre], emeta[kPre];
#pragma unroll
    for (int it = 0; it < kPre; ++it) {
        const int i = tid + it * kThreads;
        eoff[it] = 0; emeta[it] = -1;
        if (i < n_patch) {
            const int ci = i / (PH * PW);
            const int rem = i - ci * PH * PW;
            const int py = rem / PW, px = rem - py * PW;
            int c = ci, sel = 0;
            if (ci >= c0 + c1) { sel = 2; c = ci - c0 - c1; }
            else if (ci >= c0) { sel = 1; c = ci - c0; }
            eoff[it] = (c * H + (py - kPad)) * W + (px - kPad);
            emeta[it] = ((((ci * PH + py) * pl.pitch + px) * 2) << 16) | (sel << 12) | (py << 6) | px;      // py, px < 64
        }
    }
    griddep_wait();            // the input is the previous kernel's output (the sampler's x_t)
    float pre[kPre];
    // tile walk without divisions in the loop: (image, tile inside the image) advance by the grid size
    const int tiles_xy = tiles_x * tiles_y;
    const int step_b = static_cast<int>(gridDim.x) / tiles_xy, step_r = static_cast<int>(gridDim.x) - step_b * tiles_xy;
    const int tx_shift = 31 - __clz(tiles_x);
    const bool tx_pow2 = (1 << tx_shift) == tiles_x;
    int nb = static_cast<int>(blockIdx.x) / tiles_xy, nrem = static_cast<int>(blockIdx.x) - nb * tiles_xy;      // the NEXT tile to fetch
    int ny0 = 0, nx0 = 0;
    auto fetch_patch = [&]() {          // patch of tile (nb, nrem) -> registers; leaves its origin in (ny0, nx0)
        const int ty = tx_pow2 ? (nrem >> tx_shift) : nrem / tiles_x;
        ny0 = ty * TH;
        nx0 = (nrem - ty * tiles_x) * TW;
        const int o = ny0 * W + nx0;
        const float* base0 = in0 + static_cast<long long>(nb) * (c0 * H * W) + o;
        const float* base1 = c1 ? in1 + static_cast<long long>(nb) * (c1 * H * W) + o : base0;
        const float* base2 = c2 ? in2 + static_cast<long long>(nb) * (c2 * H * W) + o : base0;
#pragma unroll
        for (int it = 0; it < kPre; ++it) {
            const int m = emeta[it];
            const int y = ny0 + ((m >> 6) & 63) - kPad, x = nx0 + (m & 63) - kPad;
            const bool in = m >= 0 && static_cast<unsigned>(y) < static_cast<unsigned>(H) && static_cast<unsigned>(x) < static_cast<unsigned>(W);
            const int sel = (m >> 12) & 3;
            const float* bp = sel == 0 ? base0 : (sel == 1 ? base1 : base2);
            pre[it] = in ? __ldg(bp + eoff[it]) : 0.0f;
        }
    };
    auto advance = [&]() {
        nb += step_b; nrem += step_r;
        if (nrem >= tiles_xy) { nrem -= tiles_xy; ++nb; }
    };
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint64_t desc0 = umma_desc_sw128(0);
    auto desc = [&](uint32_t addr) -> uint64_t { return desc0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
    const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(Cout));

    // A assembly: thread = pixel pair (2pp, 2pp + 1) of the tile x units j = jq, jq + 4, ..  The patch is bf16; the eight values
    // p0..p7 starting at the even pixel are four aligned words: the even pixel's unit is (p0p1)(p2p3)(p4p5)(p6 0), the odd
    // pixel's (p1p2)(p3p4)(p5p6)(p7 0).  Accumulator row of pixel p = (p >> 1) | ((p & 1) << 6): the even pixels of the tile are
    // rows 0..63, the odd ones rows 64..127, so that one store instruction of a warp writes 32 consecutive rows (all eight
    // swizzle phases, no bank conflicts); the epilogue undoes the permutation when it stages its row.
    const int pp = tid & 63, jq = tid >> 6;
    const int r0 = 2 * pp;
    const uint32_t prow = sb + pl.patch_off + (((r0 >> tw_shift) * pl.pitch + (r0 & (TW - 1))) << 1);
    const uint32_t arow0 = sb + pl.a_off + pp * 128, arow1 = arow0 + 64 * 128;
    const uint32_t sw0 = static_cast<uint32_t>(pp & 7), sw1 = sw0;
    const uint32_t joff_a = sb + pl.joff_off;
    // epilogue geometry: thread = (accumulator row = tile pixel, column part)
    const int erow = (warp & 3) * 32 + lane, epart = warp >> 2;
    const int ecols = Cout >> 1;                        // 16, 32 or 64 columns per thread
    const int upr = Cout >> 3, upr_shift = 31 - __clz(upr);           // 16-byte units per staged output row (4, 8 or 16)
    const uint32_t abase = sb + pl.a_off;

    auto copy_out = [&](int b, int y0, int x0) {           // staged tile -> global: consecutive threads, consecutive 16 bytes of a pixel's row
        __nv_bfloat16* obase = out + ((static_cast<long long>(b) * H + y0) * W + x0) * Cout;
        for (int i = tid; i < (128 << upr_shift); i += kThreads) {
            const int pix = i >> upr_shift, u = i & (upr - 1);
            const int dy = pix >> tw_shift, dx = pix & (TW - 1);
            if (y0 + dy < H && x0 + dx < W) {
                const uint4 val = lds_128u(sb + pl.stg_off + pix * (Cout * 2) + ((u ^ (pix & (upr - 1) & 7)) << 4));
                *reinterpret_cast<uint4*>(obase + (dy * W + dx) * Cout + u * 8) = val;
            }
        }
    };
    int it = 0, pb = 0, py0 = 0, px0 = 0;
    bool have_prev = false;
    if (static_cast<int>(blockIdx.x) < total_tiles) fetch_patch();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int b = nb, y0 = ny0, x0 = nx0;      // this tile (fetched during the previous iteration)
#pragma unroll
        for (int k = 0; k < kPre; ++k)
            if (emeta[k] >= 0) sts_u16(sb + pl.patch_off + (static_cast<uint32_t>(emeta[k]) >> 16), __bfloat16_as_ushort(__float2bfloat16_rn(pre[k])));
        advance();
        if (tile + static_cast<int>(gridDim.x) < total_tiles) fetch_patch();
        __syncthreads();       // patch complete
        for (int j = jq; j < nJ; j += kThreads / 64) {
            const uint32_t pa = prow + lds_u32(joff_a + j * 4);
            const uint32_t w0 = lds_u32(pa), w1 = lds_u32(pa + 4), w2 = lds_u32(pa + 8), w3 = lds_u32(pa + 12);
            const uint32_t acol = static_cast<uint32_t>((j >> 3) * 16384), u = static_cast<uint32_t>(j & 7);
            sts_128u(arow0 + acol + ((u ^ sw0) << 4), w0, w1, w2, w3 & 0xFFFFu);
            sts_128u(arow1 + acol + ((u ^ sw1) << 4), __funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16), w3 >> 16);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int a = 0; a < pl.atoms; ++a) {
                const uint64_t ad = desc(abase + a * 16384);
                const uint64_t bd = desc(sb + pl.w_off + a * (Cout * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2u * k, bd + 2u * k, idesc, (a | k) ? 1u : 0u);
            }
            umma_commit(&bars->mma_done);
        }
        if (have_prev) copy_out(pb, py0, px0);      // the previous tile leaves while this tile's MMAs run
        if (tid == 0) mbar_wait(&bars->mma_done, static_cast<uint32_t>(it) & 1u);   // the issuer alone polls; the block barrier tells the rest
        __syncthreads();       // accumulator complete; staging buffer copied out by everyone
        tc_fence_after();
        {
            const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(epart * ecols);
            const int epix = ((erow & 63) << 1) | (erow >> 6);        // tile pixel of this accumulator row
            const uint32_t srow = sb + pl.stg_off + epix * (Cout * 2);
            const int sx = epix & (upr - 1) & 7;                      // staged rows are unit-swizzled by the pixel index
            for (int c16 = 0; c16 < (ecols >> 4); ++c16) {
                uint32_t v[16];
                tmem_ld16(trow + c16 * 16, v);
                const uint32_t ba = sb + pl.bias_off + static_cast<uint32_t>(epart * ecols + c16 * 16) * 4u;
                tmem_ld_wait();
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 bb = lds_128u(ba + i * 16);
                    w[2 * i] = pack2(__uint_as_float(v[4 * i]) + __uint_as_float(bb.x), __uint_as_float(v[4 * i + 1]) + __uint_as_float(bb.y));
                    w[2 * i + 1] = pack2(__uint_as_float(v[4 * i + 2]) + __uint_as_float(bb.z), __uint_as_float(v[4 * i + 3]) + __uint_as_float(bb.w));
                }
                const int u0 = (epart * ecols + c16 * 16) >> 3;
                sts_128u(srow + ((u0 ^ sx) << 4), w[0], w[1], w[2], w[3]);
                sts_128u(srow + (((u0 + 1) ^ sx) << 4), w[4], w[5], w[6], w[7]);
            }
        }
        tc_fence_before();     // (the next iteration's two block barriers order these reads / staging writes before its MMAs / copy-out)
        have_prev = true; pb = b; py0 = y0; px0 = x0;
    }
    __syncthreads();
    if (have_prev) copy_out(pb, py0, px0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
}

}  // namespace

bool stem_umma_supported(int Cin, int Cout, int ks, int H, int W) {
    (void)H;
    if (!(ks == kKs && Cin >= 1 && Cin <= 4 && (Cout == 32 || Cout == 64 || Cout == 128) && W >= 1)) return false;
    return stem_plan(Cin, Cout, W >= 32 ? 32 : 16).total + 1024 <= 113 * 1024;        // two CTAs per SM
}

int stem_umma_prepare_attributes() {
    return static_cast<int>(cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void launch_stem_umma(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                      void* out, int B, int H, int W, int Cout, int num_sms, cudaStream_t s, bool pdl) {
    const int Cin = c0 + c1 + c2;
    const int TW = W >= 32 ? 32 : 16, tw_shift = W >= 32 ? 5 : 4;
    const int TH = 128 / TW;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const int total = tiles_x * tiles_y * B;
    const StemPlan pl = stem_plan(Cin, Cout, TW);
    int tmem_cols = 32;
    while (tmem_cols < Cout) tmem_cols *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(total < 2 * num_sms ? total : 2 * num_sms);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.total + 1024;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, stem_umma_kernel, in0, c0, in1, c1, in2, c2, w, b, reinterpret_cast<__nv_bfloat16*>(out), H, W, Cout, TW,
                       tw_shift, tiles_x, tiles_y, total, tmem_cols);
}

}  // namespace ddm
