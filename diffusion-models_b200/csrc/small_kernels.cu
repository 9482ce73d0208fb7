// Memory-bound kernels around the tensor-core convolutions: stem conv, time-conditioning MLPs, row norms,
// the fused per-timestep sampler update and a Philox N(0,1) generator.  All are HBM/L2-bound; the rules that
// matter are coalescing, 16-byte vector access and enough CTAs to cover 148 SMs.
#include "kernels.cuh"

#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {

namespace {

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return v / (1.0f + expf(-v));                               // SiLU
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));   // exact-erf GELU
    return v;
}

// ------------------------------------------------------------------------------------------------ stem conv
// One CTA = 8 x 32 output pixels of one image, one thread per pixel, 16 output channels at a time.
constexpr int kStemTH = 8, kStemTW = 32;

__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ in0, int c0, const float* __restrict__ in1, int c1,
                 const float* __restrict__ in2, int c2, const float* __restrict__ weight,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int H, int W, int Cout,
                 int ks) {
    extern __shared__ float stem_smem[];
    const int Cin = c0 + c1 + c2;
    const int pad = ks / 2;
    const int PH = kStemTH + ks - 1, PW = kStemTW + ks - 1;
    float* w_s = stem_smem;                        // [ks*ks*Cin][Cout]
    float* patch = w_s + ks * ks * Cin * Cout;     // [Cin][PH][PW]
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kStemTH, x0 = blockIdx.x * kStemTW;
    for (int i = threadIdx.x; i < ks * ks * Cin * Cout; i += blockDim.x) w_s[i] = __ldg(weight + i);
    for (int i = threadIdx.x; i < Cin * PH * PW; i += blockDim.x) {
        const int ci = i / (PH * PW);
        const int rem = i - ci * PH * PW;
        const int py = rem / PW, px = rem - py * PW;
        const int y = y0 + py - pad, x = x0 + px - pad;
        float v = 0.0f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const float* src;
            int c, cn;
            if (ci < c0) { src = in0; c = ci; cn = c0; }
            else if (ci < c0 + c1) { src = in1; c = ci - c0; cn = c1; }
            else { src = in2; c = ci - c0 - c1; cn = c2; }
            v = __ldg(src + ((static_cast<long long>(b) * cn + c) * H + y) * W + x);
        }
        patch[i] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / kStemTW, tx = threadIdx.x % kStemTW;
    const int y = y0 + ty, x = x0 + tx;
    const bool valid = (y < H) && (x < W);
    for (int cb = 0; cb < Cout; cb += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = (cb + j < Cout) ? __ldg(bias + cb + j) : 0.0f;
        // weight rows are tap-major: k = (ky*ks + kx)*Cin + ci
        for (int ky = 0; ky < ks; ++ky) {
            for (int kx = 0; kx < ks; ++kx) {
                for (int ci = 0; ci < Cin; ++ci) {
                    const float v = patch[(ci * PH + ty + ky) * PW + tx + kx];
                    const float* wr = w_s + ((ky * ks + kx) * Cin + ci) * Cout + cb;
                    if (cb + 16 <= Cout) {
                        const float4* w4 = reinterpret_cast<const float4*>(wr);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w = w4[j];
                            acc[4 * j + 0] = fmaf(v, w.x, acc[4 * j + 0]);
                            acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
                            acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
                            acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
                        }
                    } else {
                        for (int j = 0; j < 16 && cb + j < Cout; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
                    }
                }
            }
        }
        if (valid) {
            __nv_bfloat16* o = out + ((static_cast<long long>(b) * H + y) * W + x) * Cout + cb;
            if (cb + 16 <= Cout && (Cout % 8) == 0) {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(acc[2 * j], acc[2 * j + 1]);
                reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
                for (int j = 0; j < 16 && cb + j < Cout; ++j) o[j] = __float2bfloat16_rn(acc[j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ time path
__global__ void sinusoidal_kernel(const float* __restrict__ t, float* __restrict__ out, int rows, int dim, float theta) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * half) return;
    const int r = i / half, j = i - r * half;
    const float step = logf(theta) / static_cast<float>(half - 1);
    const float f = expf(static_cast<float>(j) * -step);
    const float a = t[r] * f;
    out[r * dim + j] = sinf(a);
    out[r * dim + half + j] = cosf(a);
}

// one warp per output element (r, n); lanes stride over K (coalesced reads of W rows)
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ b,
                    float* __restrict__ y, int ldy, int rows, int N, int K, int act_in, int act_out) {
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= static_cast<long long>(rows) * N) return;
    const int r = static_cast<int>(warp / N), n = static_cast<int>(warp - static_cast<long long>(r) * N);
    const float* xr = x + static_cast<long long>(r) * ldx;
    const float* wr = W + static_cast<long long>(n) * K;
    float acc = 0.0f;
    for (int k = lane; k < K; k += 32) acc = fmaf(apply_act(xr[k], act_in), __ldg(wr + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[static_cast<long long>(r) * ldy + n] = apply_act(acc + (b != nullptr ? b[n] : 0.0f), act_out);
}

// ------------------------------------------------------------------------------------------------ row norms
// A row of C bf16 channels is covered by `lpr` lanes (power of two <= 32), 8 channels (16 B) per lane per pass.
__device__ __forceinline__ int lanes_per_row(int C) {
    int l = 1;
    while (l < 32 && l * 8 < C) l <<= 1;
    return l;
}

__global__ void __launch_bounds__(256)
row_rnorm_kernel(const __nv_bfloat16* __restrict__ x, int ld, float* __restrict__ rnorm, long long rows, int C) {
    const int lpr = lanes_per_row(C);
    const int rows_per_warp = 32 / lpr;
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long row = warp * rows_per_warp + lane / lpr;
    const int sub = lane % lpr;
    float s = 0.0f;
    if (row < rows) {
        const __nv_bfloat16* xr = x + row * ld;
        for (int c = sub * 8; c < C; c += lpr * 8) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + c));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_lo(w[j]), b2 = bf16_hi(w[j]);
                s = fmaf(a, a, fmaf(b2, b2, s));
            }
        }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < rows && sub == 0) rnorm[row] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

__global__ void __launch_bounds__(256)
rmsnorm_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ g, const float* __restrict__ ss,
                   long long ss_stride, long long rows_per_batch, int act, const __nv_bfloat16* __restrict__ res,
                   __nv_bfloat16* __restrict__ out, long long rows, int C) {
    const int lpr = lanes_per_row(C);
    const int rows_per_warp = 32 / lpr;
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long row = warp * rows_per_warp + lane / lpr;
    const int sub = lane % lpr;
    const bool live = row < rows;
    const __nv_bfloat16* xr = x + (live ? row : 0) * C;
    float s = 0.0f;
    // the lane's pieces of the row stay in registers between the two passes (up to 4 x 16 bytes: C <= 1024 at 32 lanes)
    constexpr int kKeep = 4;
    uint4 keep[kKeep];
    auto sumsq = [&](const uint4& u) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = bf16_lo(w[j]), b2 = bf16_hi(w[j]);
            s = fmaf(a, a, fmaf(b2, b2, s));
        }
    };
    if (live) {
#pragma unroll
        for (int i = 0; i < kKeep; ++i) {
            const int c = (sub + i * lpr) * 8;
            if (c < C) { keep[i] = __ldg(reinterpret_cast<const uint4*>(xr + c)); sumsq(keep[i]); }
        }
        for (int c = (sub + kKeep * lpr) * 8; c < C; c += lpr * 8) sumsq(__ldg(reinterpret_cast<const uint4*>(xr + c)));
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (!live) return;
    const float rinv = (g != nullptr) ? 1.0f / fmaxf(sqrtf(s), 1e-12f) : 1.0f;
    const float* ssr = (ss != nullptr) ? ss + (row / rows_per_batch) * ss_stride : nullptr;
    // vector loads of the per-channel parameters need 16-byte aligned rows (C % 8 == 0 is given; check the bases)
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(ssr) | (static_cast<uintptr_t>(C) * 4)) & 15) == 0;
    auto finish = [&](int c, const uint4& u) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
        float gv[8], sc[8], sh[8];
        if (vec_ok) {
            if (g != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(g + c)), b2 = __ldg(reinterpret_cast<const float4*>(g + c + 4));
                gv[0] = a.x; gv[1] = a.y; gv[2] = a.z; gv[3] = a.w; gv[4] = b2.x; gv[5] = b2.y; gv[6] = b2.z; gv[7] = b2.w;
            }
            if (ssr != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(ssr + c)), b2 = __ldg(reinterpret_cast<const float4*>(ssr + c + 4));
                const float4 d = __ldg(reinterpret_cast<const float4*>(ssr + C + c)), e = __ldg(reinterpret_cast<const float4*>(ssr + C + c + 4));
                sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b2.x; sc[5] = b2.y; sc[6] = b2.z; sc[7] = b2.w;
                sh[0] = d.x; sh[1] = d.y; sh[2] = d.z; sh[3] = d.w; sh[4] = e.x; sh[5] = e.y; sh[6] = e.z; sh[7] = e.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (g != nullptr) gv[j] = __ldg(g + c + j);
                if (ssr != nullptr) { sc[j] = __ldg(ssr + c + j); sh[j] = __ldg(ssr + C + c + j); }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = f[j];
            if (g != nullptr) t = t * rinv * gv[j];
            if (ssr != nullptr) t = fmaf(t, sc[j] + 1.0f, sh[j]);
            if (act == 1) t = __fdividef(t, 1.0f + __expf(-t));
            f[j] = t;
        }
        if (res != nullptr) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + row * C + c));
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { f[2 * j] += bf16_lo(rw[j]); f[2 * j + 1] += bf16_hi(rw[j]); }
        }
        *reinterpret_cast<uint4*>(out + row * C + c) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    };
#pragma unroll
    for (int i = 0; i < kKeep; ++i) {
        const int c = (sub + i * lpr) * 8;
        if (c < C) finish(c, keep[i]);
    }
    for (int c = (sub + kKeep * lpr) * 8; c < C; c += lpr * 8) finish(c, __ldg(reinterpret_cast<const uint4*>(xr + c)));
}

// The Block tail over split-K partial sums (conv_tc.cuh: ksplit): x[row][c] = bias[c] + sum_i part[i][row][c], then exactly
// rmsnorm_act_kernel's math.  fp32 partials [ksplit][rows][C]; the lane's pieces of the row stay in registers (C <= 1024).
__global__ void __launch_bounds__(256)
rmsnorm_act_split_kernel(const float* __restrict__ part, int ksplit, const float* __restrict__ bias, const float* __restrict__ g,
                         const float* __restrict__ ss, long long ss_stride, long long rows_per_batch, int act,
                         const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ out, long long rows, int C) {
    const int lpr = lanes_per_row(C);
    const int rows_per_warp = 32 / lpr;
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long row = warp * rows_per_warp + lane / lpr;
    const int sub = lane % lpr;
    const bool live = row < rows;
    constexpr int kKeep = 4;
    float keep[kKeep][8];
    float s = 0.0f;
    if (live) {
#pragma unroll
        for (int i = 0; i < kKeep; ++i) {
            const int c = (sub + i * lpr) * 8;
            if (c < C) {
                float4 a = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 b = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < ksplit; ++k) {          // fixed order: the sum does not depend on the launch geometry
                    const float* pr = part + (static_cast<long long>(k) * rows + row) * C + c;
                    const float4 u = __ldg(reinterpret_cast<const float4*>(pr)), v = __ldg(reinterpret_cast<const float4*>(pr + 4));
                    a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
                    b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
                }
                keep[i][0] = a.x; keep[i][1] = a.y; keep[i][2] = a.z; keep[i][3] = a.w;
                keep[i][4] = b.x; keep[i][5] = b.y; keep[i][6] = b.z; keep[i][7] = b.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(keep[i][j], keep[i][j], s);
            }
        }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (!live) return;
    const float rinv = (g != nullptr) ? 1.0f / fmaxf(sqrtf(s), 1e-12f) : 1.0f;
    const float* ssr = (ss != nullptr) ? ss + (row / rows_per_batch) * ss_stride : nullptr;
#pragma unroll
    for (int i = 0; i < kKeep; ++i) {
        const int c = (sub + i * lpr) * 8;
        if (c >= C) continue;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = keep[i][j];
            if (g != nullptr) t = t * rinv * __ldg(g + c + j);
            if (ssr != nullptr) t = fmaf(t, __ldg(ssr + c + j) + 1.0f, __ldg(ssr + C + c + j));
            if (act == 1) t = __fdividef(t, 1.0f + __expf(-t));
            f[j] = t;
        }
        if (res != nullptr) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + row * C + c));
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { f[2 * j] += bf16_lo(rw[j]); f[2 * j + 1] += bf16_hi(rw[j]); }
        }
        *reinterpret_cast<uint4*>(out + row * C + c) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm (+ swish)
// VAE decoder norm (ldm/modules/diffusionmodules/model.py:55-56 Normalize = GroupNorm(32, C, eps=1e-6, affine) and the
// x * sigmoid(x) that follows it at :118-119, :127-128, :573-574).  The statistics span a whole image (H*W x C/G values per
// group), so this is a pre-pass in front of the tcgen05 conv rather than part of its epilogue: one CTA per image, two sweeps
// over its bf16 channels-last activations (the second one from L2).
__global__ void __launch_bounds__(512)
groupnorm_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ out, int HW, int C, int G, float eps, int act) {
    extern __shared__ float gsm[];            // [C] sum | [C] sum of squares -> reused as [C] scale | [C] shift ; [G] mean | [G] rstd ;
    float* ch_a = gsm;                        // then per-thread partial sums [512][16] (fixed-order reduction: bitwise repeatable
    float* ch_b = gsm + C;                    // and independent of the batch, unlike shared-memory atomics)
    float* g_mean = gsm + 2 * C;
    float* g_rstd = g_mean + G;
    float* part = g_rstd + G;
    const int tid = threadIdx.x;
    const __nv_bfloat16* xb = x + static_cast<long long>(blockIdx.x) * HW * C;
    __nv_bfloat16* ob = out + static_cast<long long>(blockIdx.x) * HW * C;
    const int nvec = C >> 3;                  // 16-byte vectors per pixel
    const int rstep = blockDim.x / nvec;      // pixels in flight per sweep step
    const int v = tid % nvec, r0 = tid / nvec;
    const bool live = r0 < rstep;
    {
        float s[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = 0.0f; q[j] = 0.0f; }
        for (int row = r0; live && row < HW; row += rstep) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(row) * C) + v);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_lo(w[j]), b2 = bf16_hi(w[j]);
                s[2 * j] += a; s[2 * j + 1] += b2;
                q[2 * j] = fmaf(a, a, q[2 * j]); q[2 * j + 1] = fmaf(b2, b2, q[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { part[tid * 16 + j] = s[j]; part[tid * 16 + 8 + j] = q[j]; }
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {       // channel c = vector c / 8, element c % 8: the rstep pixel-lanes in order
        float s = 0.0f, q = 0.0f;
        for (int r = 0; r < rstep; ++r) {
            const float* pp = part + ((r * nvec + (c >> 3)) * 16) + (c & 7);
            s += pp[0];
            q += pp[8];
        }
        ch_a[c] = s;
        ch_b[c] = q;
    }
    __syncthreads();
    const int cg = C / G;
    if (tid < G) {
        float s = 0.0f, q = 0.0f;
        for (int c = tid * cg; c < (tid + 1) * cg; ++c) { s += ch_a[c]; q += ch_b[c]; }
        const float inv_n = 1.0f / (static_cast<float>(HW) * static_cast<float>(cg));
        const float mean = s * inv_n;
        const float var = fmaxf(q * inv_n - mean * mean, 0.0f);          // biased variance, like torch.nn.GroupNorm
        g_mean[tid] = mean;
        g_rstd[tid] = rsqrtf(var + eps);
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        const int g = c / cg;
        const float sc = __ldg(gamma + c) * g_rstd[g];
        ch_a[c] = sc;
        ch_b[c] = __ldg(beta + c) - g_mean[g] * sc;
    }
    __syncthreads();
    if (!live) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = ch_a[v * 8 + j]; sh[j] = ch_b[v * 8 + j]; }
    for (int row = r0; row < HW; row += rstep) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(row) * C) + v);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = fmaf(f[j], sc[j], sh[j]);
            if (act == 1) t = __fdividef(t, 1.0f + __expf(-t));
            f[j] = t;
        }
        *(reinterpret_cast<uint4*>(ob + static_cast<long long>(row) * C) + v) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
}

// ------------------------------------------------------------------------------------------------ head conv
// final_conv (1x1, C -> N <= 8, denoising_diffusion.py:343,390): bf16 channels-last rows -> fp32 NCHW planes.
// HBM-bound (reads 2*C bytes per pixel): `lpr` lanes share a pixel with 16-byte loads, shuffle-reduce, coalesced plane
// stores.  fp32 weights (no rounding of the last layer's weights).
template <int NOUT>
__global__ void __launch_bounds__(256)
head_conv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out, long long rows, int C, int HW) {
    extern __shared__ float head_w[];                // [NOUT][C]
    for (int i = threadIdx.x; i < NOUT * C; i += blockDim.x) head_w[i] = __ldg(w + i);
    __syncthreads();
    // one thread per pixel: C/8 independent 16-byte loads in flight per thread, plane stores coalesced across the warp
    const long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float acc[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[n] = __ldg(bias + n);
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
    for (int c8 = 0; c8 < C / 8; c8 += 4) {
        uint4 u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (c8 + i < C / 8) ? __ldg(xr + c8 + i) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (c8 + i < C / 8) {
                const uint32_t ww[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
                float f[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(ww[j]); f[2 * j + 1] = bf16_hi(ww[j]); }
                const int c = (c8 + i) * 8;
#pragma unroll
                for (int n = 0; n < NOUT; ++n) {
                    const float4 w0 = *reinterpret_cast<const float4*>(head_w + n * C + c);
                    const float4 w1 = *reinterpret_cast<const float4*>(head_w + n * C + c + 4);
                    acc[n] = fmaf(f[0], w0.x, fmaf(f[1], w0.y, fmaf(f[2], w0.z, fmaf(f[3], w0.w, acc[n]))));
                    acc[n] = fmaf(f[4], w1.x, fmaf(f[5], w1.y, fmaf(f[6], w1.z, fmaf(f[7], w1.w, acc[n]))));
                }
            }
        }
    }
    const long long b = row / HW, pix = row - b * HW;
#pragma unroll
    for (int n = 0; n < NOUT; ++n) out[(b * NOUT + n) * HW + pix] = acc[n];
}

// ------------------------------------------------------------------------------------------------ Philox N(0,1)
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return {c0, c1, c2, c3};
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u = (static_cast<float>(a) + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
    const float v = (static_cast<float>(b) + 0.5f) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * logf(u));
    float s, c;
    sincospif(2.0f * v, &s, &c);
    n0 = r * c;
    n1 = r * s;
}
// four N(0,1) values for the element group `grp` (= element index / 4) of stream (seed, sid)
__device__ __forceinline__ void normal4(unsigned long long seed, unsigned long long sid, unsigned long long grp, float (&z)[4]) {
    const Philox4 p = philox4x32_10(static_cast<uint32_t>(grp), static_cast<uint32_t>(grp >> 32), static_cast<uint32_t>(sid),
                                    static_cast<uint32_t>(sid >> 32), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    box_muller(p.x, p.y, z[0], z[1]);
    box_muller(p.z, p.w, z[2], z[3]);
}

__global__ void __launch_bounds__(256)
randn_kernel(float* __restrict__ x, unsigned long long seed, unsigned long long sid, long long numel) {
    const long long grp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i = grp * 4;
    if (i >= numel) return;
    float z[4];
    normal4(seed, sid, static_cast<unsigned long long>(grp), z);
    for (int j = 0; j < 4 && i + j < numel; ++j) x[i + j] = z[j];
}

// ------------------------------------------------------------------------------------------------ sampler step
// Separate mul/sub/div roundings (no FMA contraction) so that the fp32 update matches the reference's chain of
// elementwise ATen ops bit for bit (denoising_diffusion.py:570-580, 596-601, 699-701).
__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

__global__ void __launch_bounds__(256)
sampler_step_kernel(int kind, float* __restrict__ x, const float* __restrict__ mo, const float* __restrict__ noise_base,
                    long long noise_stride, float* __restrict__ x0_out, const float* __restrict__ coef_tab, int* __restrict__ step_counter,
                    int objective, unsigned long long seed, long long numel) {
    const int step = *step_counter;
    const float* noise = noise_base != nullptr ? noise_base + static_cast<long long>(step) * noise_stride : nullptr;
    const float* cf = coef_tab + static_cast<long long>(step) * 8;
    const float ra = cf[0], rm1 = cf[1], k2 = cf[2], k3 = cf[3], k4 = cf[4], k5 = cf[5], sac = cf[6], s1m = cf[7];
    const long long grp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i0 = grp * 4;
    if (i0 >= numel) return;
    const float noise_amp = k4;                       // sigma (DDIM) or exp(0.5 logvar) (DDPM; 0 at t == 0)
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise_amp != 0.0f && noise == nullptr)   // stream id = (call epoch, step): a replayed graph draws fresh noise every call
        normal4(seed, (static_cast<unsigned long long>(static_cast<unsigned>(step_counter[1])) << 32) | (static_cast<unsigned long long>(step) + 1ull),
                static_cast<unsigned long long>(grp), z);
    for (int j = 0; j < 4 && i0 + j < numel; ++j) {
        const long long i = i0 + j;
        const float xt = x[i], o = mo[i];
        const float rax = __fmul_rn(ra, xt);
        float x0, eps;
        if (objective == 0) {
            x0 = clamp1(__fsub_rn(rax, __fmul_rn(rm1, o)));
            eps = (kind == DDM_KIND_DDIM) ? __fdiv_rn(__fsub_rn(rax, x0), rm1) : o;
        } else {
            x0 = (objective == 1) ? clamp1(o) : clamp1(__fsub_rn(__fmul_rn(sac, xt), __fmul_rn(s1m, o)));
            eps = __fdiv_rn(__fsub_rn(rax, x0), rm1);
        }
        const float zi = (noise != nullptr && noise_amp != 0.0f) ? noise[i] : z[j];
        float xn;
        if (kind == DDM_KIND_DDIM) {
            if (k5 != 0.0f) {
                xn = x0;                                                                  // t_next < 0 (dd:686-689)
            } else {
                xn = __fadd_rn(__fmul_rn(x0, k2), __fmul_rn(k3, eps));                      // x0*sqrt(a_next) + c*eps
                if (noise_amp != 0.0f) xn = __fadd_rn(xn, __fmul_rn(noise_amp, zi));      // + sigma*noise
            }
        } else {
            xn = __fadd_rn(__fmul_rn(k2, x0), __fmul_rn(k3, xt));                           // posterior mean (dd:596-598)
            if (noise_amp != 0.0f) xn = __fadd_rn(xn, __fmul_rn(noise_amp, zi));          // dd:644
        }
        x[i] = xn;
        if (x0_out != nullptr) x0_out[i] = x0;
    }
}

// Guided DDIM step (denoising_diffusion.py:710-777 ddim_sample_guided): model_predictions WITHOUT re-deriving the noise
// (rederive_pred_noise is left at False there), optional x0 clamp, the DDIM update, then the known region is replaced by
// the guide noised to step t:  x = x * mask + q_sample(guide, t) * (1 - mask)   (dd:747-749; not on the last step).
__global__ void __launch_bounds__(256)
sampler_step_guided_kernel(float* __restrict__ x, const float* __restrict__ mo, const float* __restrict__ noise_base, long long noise_stride,
                           const float* __restrict__ guide, const float* __restrict__ mask, const float* __restrict__ gnoise_base,
                           long long gnoise_stride, float* __restrict__ x0_out, const float* __restrict__ coef_tab,
                           int* __restrict__ step_counter, int objective, int clip, unsigned long long seed, long long numel) {
    const int step = *step_counter;
    const float* noise = noise_base != nullptr ? noise_base + static_cast<long long>(step) * noise_stride : nullptr;
    const float* gnoise = gnoise_base != nullptr ? gnoise_base + static_cast<long long>(step) * gnoise_stride : nullptr;
    const float* cf = coef_tab + static_cast<long long>(step) * 8;
    const float ra = cf[0], rm1 = cf[1], k2 = cf[2], k3 = cf[3], sigma = cf[4], last = cf[5], sac = cf[6], s1m = cf[7];
    const long long grp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i0 = grp * 4;
    if (i0 >= numel) return;
    const unsigned long long epoch = static_cast<unsigned long long>(static_cast<unsigned>(step_counter[1])) << 32;
    float z[4] = {0.f, 0.f, 0.f, 0.f}, zg[4] = {0.f, 0.f, 0.f, 0.f};
    if (sigma != 0.0f && noise == nullptr && last == 0.0f)
        normal4(seed, epoch | (static_cast<unsigned long long>(step) + 1ull), static_cast<unsigned long long>(grp), z);
    if (guide != nullptr && gnoise == nullptr && last == 0.0f)
        normal4(seed, epoch | (0x40000000ull + static_cast<unsigned long long>(step)), static_cast<unsigned long long>(grp), zg);
    for (int j = 0; j < 4 && i0 + j < numel; ++j) {
        const long long i = i0 + j;
        const float xt = x[i], o = mo[i];
        const float rax = __fmul_rn(ra, xt);
        float x0, eps;
        if (objective == 0) {
            x0 = __fsub_rn(rax, __fmul_rn(rm1, o));
            if (clip) x0 = clamp1(x0);
            eps = o;
        } else {
            x0 = (objective == 1) ? o : __fsub_rn(__fmul_rn(sac, xt), __fmul_rn(s1m, o));
            if (clip) x0 = clamp1(x0);
            eps = __fdiv_rn(__fsub_rn(rax, x0), rm1);
        }
        float xn;
        if (last != 0.0f) {
            xn = x0;
        } else {
            xn = __fadd_rn(__fmul_rn(x0, k2), __fmul_rn(k3, eps));
            if (sigma != 0.0f) xn = __fadd_rn(xn, __fmul_rn(sigma, noise != nullptr ? noise[i] : z[j]));
            if (guide != nullptr) {
                const float gt = __fadd_rn(__fmul_rn(sac, guide[i]), __fmul_rn(s1m, gnoise != nullptr ? gnoise[i] : zg[j]));
                const float m = mask[i];
                xn = __fadd_rn(__fmul_rn(xn, m), __fmul_rn(gt, __fsub_rn(1.0f, m)));
            }
        }
        x[i] = xn;
        if (x0_out != nullptr) x0_out[i] = x0;
    }
}

// Ancestral step of LearnedGaussianDiffusion (learned_gaussian_diffusion.py:91-111 + dd:638-645): the network output has
// 2C channels per sample, (pred_noise | variance interpolation fraction v in [-1, 1]):
//   logvar = f * log(beta_t) + (1 - f) * posterior_log_variance_clipped_t,  f = (v + 1) / 2
//   x_{t-1} = coef1 * clamp(x0) + coef2 * x_t + exp(0.5 * logvar) * z            (z = 0 at t = 0)
// coef row: { sqrt_recip_acp, sqrt_recipm1_acp, coef1, coef2, noise_on, min_log, max_log, - }
__global__ void __launch_bounds__(256)
sampler_step_learned_kernel(float* __restrict__ x, const float* __restrict__ mo, const float* __restrict__ noise_base,
                            long long noise_stride, float* __restrict__ x0_out, const float* __restrict__ coef_tab,
                            int* __restrict__ step_counter, unsigned long long seed, long long numel, long long per_sample) {
    const int step = *step_counter;
    const float* noise = noise_base != nullptr ? noise_base + static_cast<long long>(step) * noise_stride : nullptr;
    const float* cf = coef_tab + static_cast<long long>(step) * 8;
    const float ra = cf[0], rm1 = cf[1], c1 = cf[2], c2 = cf[3], noise_on = cf[4], min_log = cf[5], max_log = cf[6];
    const long long grp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i0 = grp * 4;
    if (i0 >= numel) return;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise_on != 0.0f && noise == nullptr)
        normal4(seed, (static_cast<unsigned long long>(static_cast<unsigned>(step_counter[1])) << 32) | (static_cast<unsigned long long>(step) + 1ull),
                static_cast<unsigned long long>(grp), z);
    for (int j = 0; j < 4 && i0 + j < numel; ++j) {
        const long long i = i0 + j;
        const long long b = i / per_sample, r = i - b * per_sample;
        const float xt = x[i];
        const float eps = mo[b * 2 * per_sample + r], v = mo[b * 2 * per_sample + per_sample + r];
        const float x0 = clamp1(__fsub_rn(__fmul_rn(ra, xt), __fmul_rn(rm1, eps)));
        float xn = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt));
        if (noise_on != 0.0f) {
            const float f = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
            const float logvar = __fadd_rn(__fmul_rn(f, max_log), __fmul_rn(__fsub_rn(1.0f, f), min_log));
            const float zi = noise != nullptr ? noise[i] : z[j];
            xn = __fadd_rn(xn, __fmul_rn(expf(__fmul_rn(0.5f, logvar)), zi));
        }
        x[i] = xn;
        if (x0_out != nullptr) x0_out[i] = x0;
    }
}

__global__ void bump_counter_kernel(int* c) { *c += 1; }

__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ x, float* __restrict__ y, int unnorm, long long numel) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    const float v = x[i];
    y[i] = unnorm ? __fmul_rn(__fadd_rn(v, 1.0f), 0.5f) : v;
}

__global__ void __launch_bounds__(256)
select_row_kernel(const float* __restrict__ table, const int* __restrict__ step_counter, float* __restrict__ dst, int row_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row_len) dst[i] = table[static_cast<long long>(*step_counter) * row_len + i];
}

inline unsigned blocks_for(long long n, int per_block) { return static_cast<unsigned>((n + per_block - 1) / per_block); }

}  // namespace

// ------------------------------------------------------------------------------------------------ launchers
int stem_smem_bytes(int Cin, int Cout, int ks) {
    return (ks * ks * Cin * Cout + Cin * (kStemTH + ks - 1) * (kStemTW + ks - 1)) * 4;
}
int stem_prepare_attributes() {
    return static_cast<int>(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}
void launch_stem(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                 void* out, int B, int H, int W, int Cout, int ks, cudaStream_t s) {
    dim3 grid((W + kStemTW - 1) / kStemTW, (H + kStemTH - 1) / kStemTH, B);
    stem_conv_kernel<<<grid, 256, stem_smem_bytes(c0 + c1 + c2, Cout, ks), s>>>(
        in0, c0, in1, c1, in2, c2, w, b, reinterpret_cast<__nv_bfloat16*>(out), B, H, W, Cout, ks);
}
void launch_sinusoidal(const float* t, float* out, int rows, int dim, float theta, cudaStream_t s) {
    const int n = rows * (dim / 2);
    sinusoidal_kernel<<<blocks_for(n, 128), 128, 0, s>>>(t, out, rows, dim, theta);
}
void launch_small_linear(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int rows, int N, int K,
                         int act_in, int act_out, cudaStream_t s) {
    const long long warps = static_cast<long long>(rows) * N;
    small_linear_kernel<<<blocks_for(warps, 8), 256, 0, s>>>(x, ldx, W, b, y, ldy, rows, N, K, act_in, act_out);
}
void launch_row_rnorm(const void* x, int ld, float* rn, long long rows, int C, cudaStream_t s) {
    int lpr = 1;
    while (lpr < 32 && lpr * 8 < C) lpr <<= 1;
    const long long warps = (rows + (32 / lpr) - 1) / (32 / lpr);
    row_rnorm_kernel<<<blocks_for(warps, 8), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rn, rows, C);
}
void launch_rmsnorm_act(const void* x, const float* g, const float* ss, long long ss_stride, long long rows_per_batch, int act,
                        const void* res, void* out, long long rows, int C, cudaStream_t s) {
    int lpr = 1;
    while (lpr < 32 && lpr * 8 < C) lpr <<= 1;
    const long long warps = (rows + (32 / lpr) - 1) / (32 / lpr);
    rmsnorm_act_kernel<<<blocks_for(warps, 8), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), g, ss, ss_stride,
                                                          rows_per_batch, act, reinterpret_cast<const __nv_bfloat16*>(res),
                                                          reinterpret_cast<__nv_bfloat16*>(out), rows, C);
}
void launch_rmsnorm_act_split(const float* partials, int ksplit, const float* bias, const float* g, const float* ss, long long ss_stride,
                              long long rows_per_batch, int act, const void* res, void* out, long long rows, int C, cudaStream_t s) {
    int lpr = 1;
    while (lpr < 32 && lpr * 8 < C) lpr <<= 1;
    const long long warps = (rows + (32 / lpr) - 1) / (32 / lpr);
    rmsnorm_act_split_kernel<<<blocks_for(warps, 8), 256, 0, s>>>(partials, ksplit, bias, g, ss, ss_stride, rows_per_batch, act,
                                                                reinterpret_cast<const __nv_bfloat16*>(res),
                                                                reinterpret_cast<__nv_bfloat16*>(out), rows, C);
}
int launch_groupnorm_act(const void* x, const float* gamma, const float* beta, void* out, int B, int HW, int C, int G, float eps, int act,
                         cudaStream_t s) {
    const int nvec = C / 8;
    if (C % 8 != 0 || nvec > 512 || G < 1 || G > 512 || C % G != 0) return -3;
    const int smem = (2 * C + 2 * G + 512 * 16) * 4;
    if (smem > 48 * 1024) return -3;
    groupnorm_act_kernel<<<B, 512, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), HW,
                                              C, G, eps, act);
    return 0;
}
int launch_head_conv(const void* x, const float* w, const float* bias, float* out, long long rows, int C, int N, int HW,
                     cudaStream_t s) {
    const auto* xb = reinterpret_cast<const __nv_bfloat16*>(x);
    const unsigned grid = blocks_for(rows, 256);
    const int smem = N * C * 4;
    switch (N) {
        case 1: head_conv_kernel<1><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        case 2: head_conv_kernel<2><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        case 3: head_conv_kernel<3><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        case 4: head_conv_kernel<4><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        case 6: head_conv_kernel<6><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        case 8: head_conv_kernel<8><<<grid, 256, smem, s>>>(xb, w, bias, out, rows, C, HW); return 0;
        default: return -3;
    }
}
void launch_sampler_step(int kind, float* x, const float* mo, const float* noise, long long noise_stride, float* x0_out, const float* coef, int* step_counter,
                         int advance, int objective, unsigned long long seed, long long numel, cudaStream_t s) {
    sampler_step_kernel<<<blocks_for((numel + 3) / 4, 256), 256, 0, s>>>(kind, x, mo, noise, noise_stride, x0_out, coef, step_counter, objective,
                                                                          seed, numel);
    if (advance) bump_counter_kernel<<<1, 1, 0, s>>>(step_counter);
}
void launch_sampler_step_learned(float* x, const float* mo, const float* noise, long long noise_stride, float* x0_out, const float* coef,
                                 int* step_counter, int advance, unsigned long long seed, long long numel, long long per_sample, cudaStream_t s) {
    sampler_step_learned_kernel<<<blocks_for((numel + 3) / 4, 256), 256, 0, s>>>(x, mo, noise, noise_stride, x0_out, coef, step_counter, seed,
                                                                                  numel, per_sample);
    if (advance) bump_counter_kernel<<<1, 1, 0, s>>>(step_counter);
}
void launch_sampler_step_guided(float* x, const float* mo, const float* noise, long long noise_stride, const float* guide, const float* mask,
                                const float* gnoise, long long gnoise_stride, float* x0_out, const float* coef, int* step_counter, int advance,
                                int objective, int clip, unsigned long long seed, long long numel, cudaStream_t s) {
    sampler_step_guided_kernel<<<blocks_for((numel + 3) / 4, 256), 256, 0, s>>>(x, mo, noise, noise_stride, guide, mask, gnoise, gnoise_stride,
                                                                                 x0_out, coef, step_counter, objective, clip, seed, numel);
    if (advance) bump_counter_kernel<<<1, 1, 0, s>>>(step_counter);
}
void launch_finalize(const float* x, float* y, int unnorm, long long numel, cudaStream_t s) {
    finalize_kernel<<<blocks_for(numel, 256), 256, 0, s>>>(x, y, unnorm, numel);
}
void launch_select_row(const float* table, const int* step_counter, float* dst, int row_len, cudaStream_t s) {
    select_row_kernel<<<blocks_for(row_len, 256), 256, 0, s>>>(table, step_counter, dst, row_len);
}
void launch_randn(float* x, unsigned long long seed, unsigned long long sid, long long numel, cudaStream_t s) {
    randn_kernel<<<blocks_for((numel + 3) / 4, 256), 256, 0, s>>>(x, seed, sid, numel);
}

}  // namespace ddm
