// Attention cores of the U-Net (everything between the to_qkv and to_out 1x1 convolutions, which run on the
// tensor-core conv kernel):
//   linear_attention_kernel : LinearAttention.forward, denoising_diffusion.py:178-192 (O(n) in the token count)
//   attention_kernel        : Attention.forward + Attend.forward, denoising_diffusion.py:220-228, attend.py:109-124,
//                             and CrossAttention.forward, denoising_diffusion_text_conditional.py:66-77
// Both are <2 % of the network FLOPs and bound by the traffic of the qkv tensor: bf16 in/out with 16-byte
// accesses, fp32 math, one CTA per (batch, head) so the 32x32 context / the K,V tiles live in shared memory.
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
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ------------------------------------------------------------------------------------------------ linear attention
template <int D>
__global__ void __launch_bounds__(256)
linear_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ mem_kv,
                        __nv_bfloat16* __restrict__ out, int n, int heads, int n_mem) {
    constexpr int P = D / 8;            // 16-byte parts per token row
    constexpr int TOK = 256 / P;        // tokens staged per pass
    constexpr int EPT = D * D / 256;    // context entries per thread
    constexpr int GPR = D / EPT;        // thread groups per context row
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int HD = heads * D, ld = 3 * HD;
    const __nv_bfloat16* base = qkv + static_cast<long long>(b) * n * ld;
    const __nv_bfloat16* qp = base + h * D;
    const __nv_bfloat16* kp = base + HD + h * D;
    const __nv_bfloat16* vp = base + 2 * HD + h * D;
    const float* mk = mem_kv + (static_cast<long long>(0) * heads + h) * D * n_mem;   // [d][n_mem]
    const float* mv = mem_kv + (static_cast<long long>(1) * heads + h) * D * n_mem;

    __shared__ float ek[TOK * D];       // exp(k - kmax) tile, also scratch for the max reduction
    __shared__ float vv[TOK * D];
    __shared__ float ctx[D * D];
    __shared__ float kmax[D], ksum[D];

    const int tok_l = tid / P, part = tid % P;

    // ---- pass 1: per-channel max of k over all tokens (softmax over n, dd:185)
    float mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
    for (int t0 = 0; t0 < n; t0 += TOK) {
        const int tok = t0 + tok_l;
        if (tok < n) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(kp + static_cast<long long>(tok) * ld) + part), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], f[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ek[tok_l * D + part * 8 + j] = mx[j];
    __syncthreads();
    if (tid < D) {
        float m = -INFINITY;
        for (int t = 0; t < TOK; ++t) m = fmaxf(m, ek[t * D + tid]);
        for (int t = 0; t < n_mem; ++t) m = fmaxf(m, __ldg(mk + tid * n_mem + t));
        kmax[tid] = m;
    }
    __syncthreads();

    // ---- pass 2: context[d][e] = sum_tok softmax_k[d][tok] * v[e][tok]   (dd:189), memory tokens first (dd:181-182)
    const int dd = tid / GPR, e0 = (tid % GPR) * EPT;
    float acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = 0.0f;
    float ks = 0.0f;
    for (int t0 = -TOK; t0 < n; t0 += TOK) {
        int cnt;
        if (t0 < 0) {
            cnt = n_mem;
            for (int i = tid; i < n_mem * D; i += 256) {
                const int t = i / D, d = i - t * D;
                ek[i] = __expf(__ldg(mk + d * n_mem + t) - kmax[d]);
                vv[i] = __ldg(mv + d * n_mem + t);
            }
        } else {
            cnt = min(TOK, n - t0);
            const int tok = t0 + tok_l;
            if (tok < n) {
                float f[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(kp + static_cast<long long>(tok) * ld) + part), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) ek[tok_l * D + part * 8 + j] = __expf(f[j] - kmax[part * 8 + j]);
                unpack8(__ldg(reinterpret_cast<const uint4*>(vp + static_cast<long long>(tok) * ld) + part), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) vv[tok_l * D + part * 8 + j] = f[j];
            }
        }
        __syncthreads();
        for (int t = 0; t < cnt; ++t) {
            const float a = ek[t * D + dd];
            ks += a;
#pragma unroll
            for (int j = 0; j < EPT; ++j) acc[j] = fmaf(a, vv[t * D + e0 + j], acc[j]);
        }
        __syncthreads();
    }
    if (e0 == 0) ksum[dd] = ks;
    __syncthreads();
    {
        const float sc = rsqrtf(static_cast<float>(D)) / ksum[dd];     // fold q's d^-0.5 scale (dd:187) into the context
#pragma unroll
        for (int j = 0; j < EPT; ++j) ctx[dd * D + e0 + j] = acc[j] * sc;
    }
    __syncthreads();

    // ---- pass 3: out[e][tok] = sum_d context[d][e] * softmax_d(q)[d][tok]   (dd:184,191)
    for (int tok = tid; tok < n; tok += 256) {
        float qv[D];
        const uint4* qrow = reinterpret_cast<const uint4*>(qp + static_cast<long long>(tok) * ld);
#pragma unroll
        for (int c = 0; c < P; ++c) {
            float f[8];
            unpack8(__ldg(qrow + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[c * 8 + j] = f[j];
        }
        float m = qv[0];
#pragma unroll
        for (int d = 1; d < D; ++d) m = fmaxf(m, qv[d]);
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) { qv[d] = __expf(qv[d] - m); s += qv[d]; }
        const float inv = 1.0f / s;
        float o[D];
#pragma unroll
        for (int e = 0; e < D; ++e) o[e] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float w = qv[d] * inv;
            const float4* cr = reinterpret_cast<const float4*>(ctx + d * D);
#pragma unroll
            for (int e4 = 0; e4 < D / 4; ++e4) {
                const float4 c = cr[e4];
                o[4 * e4 + 0] = fmaf(w, c.x, o[4 * e4 + 0]);
                o[4 * e4 + 1] = fmaf(w, c.y, o[4 * e4 + 1]);
                o[4 * e4 + 2] = fmaf(w, c.z, o[4 * e4 + 2]);
                o[4 * e4 + 3] = fmaf(w, c.w, o[4 * e4 + 3]);
            }
        }
        uint4* orow = reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * n + tok) * HD + h * D);
#pragma unroll
        for (int c = 0; c < P; ++c)
            orow[c] = make_uint4(pack_bf16x2(o[c * 8 + 0], o[c * 8 + 1]), pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]),
                                 pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]), pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]));
    }
}

// ------------------------------------------------------------------------------------------------ softmax attention
// One thread per query row, online softmax over key/value tiles staged in shared memory as fp32.
template <int D>
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                 const __nv_bfloat16* __restrict__ v, int ldv, const float* __restrict__ mem_k,
                 const float* __restrict__ mem_v, int n_mem, __nv_bfloat16* __restrict__ out, int nq, int nk, int heads) {
    constexpr int KT = 64;
    constexpr int P = D / 8;
    __shared__ float Ks[KT * D];
    __shared__ float Vs[KT * D];
    const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
    const int qi = blockIdx.x * blockDim.x + tid;
    const bool active = qi < nq;
    const float qscale = rsqrtf(static_cast<float>(D)) * 1.4426950408889634f;   // d^-0.5 * log2(e)

    float qv[D], acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { qv[d] = 0.0f; acc[d] = 0.0f; }
    if (active) {
        const uint4* qrow = reinterpret_cast<const uint4*>(q + (static_cast<long long>(b) * nq + qi) * ldq + h * D);
#pragma unroll
        for (int c = 0; c < P; ++c) {
            float f[8];
            unpack8(__ldg(qrow + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[c * 8 + j] = f[j] * qscale;
        }
    }
    float m = -INFINITY, l = 0.0f;

    for (int t0 = (n_mem > 0 ? -KT : 0); t0 < nk; t0 += KT) {
        int cnt;
        if (t0 < 0) {                                            // learned memory keys/values come first (dd:223-224)
            cnt = n_mem;
            for (int i = tid; i < n_mem * D; i += blockDim.x) {
                Ks[i] = __ldg(mem_k + static_cast<long long>(h) * n_mem * D + i);
                Vs[i] = __ldg(mem_v + static_cast<long long>(h) * n_mem * D + i);
            }
        } else {
            cnt = min(KT, nk - t0);
            for (int i = tid; i < cnt * P; i += blockDim.x) {
                const int t = i / P, c = i - t * P;
                const long long row = static_cast<long long>(b) * nk + t0 + t;
                float f[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(k + row * ldk + h * D) + c), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) Ks[t * D + c * 8 + j] = f[j];
                unpack8(__ldg(reinterpret_cast<const uint4*>(v + row * ldv + h * D) + c), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) Vs[t * D + c * 8 + j] = f[j];
            }
        }
        __syncthreads();
        if (active) {
            for (int t = 0; t < cnt; ++t) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + t * D);
                float s = 0.0f;
#pragma unroll
                for (int d4 = 0; d4 < D / 4; ++d4) {
                    const float4 kk = kr[d4];
                    s = fmaf(qv[4 * d4 + 0], kk.x, s);
                    s = fmaf(qv[4 * d4 + 1], kk.y, s);
                    s = fmaf(qv[4 * d4 + 2], kk.z, s);
                    s = fmaf(qv[4 * d4 + 3], kk.w, s);
                }
                if (s > m) {
                    const float corr = exp2f(m - s);
                    l *= corr;
#pragma unroll
                    for (int d = 0; d < D; ++d) acc[d] *= corr;
                    m = s;
                }
                const float pexp = exp2f(s - m);
                l += pexp;
                const float4* vr = reinterpret_cast<const float4*>(Vs + t * D);
#pragma unroll
                for (int d4 = 0; d4 < D / 4; ++d4) {
                    const float4 vw = vr[d4];
                    acc[4 * d4 + 0] = fmaf(pexp, vw.x, acc[4 * d4 + 0]);
                    acc[4 * d4 + 1] = fmaf(pexp, vw.y, acc[4 * d4 + 1]);
                    acc[4 * d4 + 2] = fmaf(pexp, vw.z, acc[4 * d4 + 2]);
                    acc[4 * d4 + 3] = fmaf(pexp, vw.w, acc[4 * d4 + 3]);
                }
            }
        }
        __syncthreads();
    }
    if (active) {
        const float inv = 1.0f / l;
        uint4* orow = reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * nq + qi) * (heads * D) + h * D);
#pragma unroll
        for (int c = 0; c < P; ++c)
            orow[c] = make_uint4(pack_bf16x2(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv),
                                 pack_bf16x2(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv),
                                 pack_bf16x2(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv),
                                 pack_bf16x2(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv));
    }
}

}  // namespace

int attention_prepare_attributes() { return 0; }

int launch_linear_attention(const void* qkv, const float* mem_kv, const float* k_shift, void* out, int B, int n, int heads, int d,
                            int n_mem, cudaStream_t s) {
    const dim3 grid(heads, B);
    const auto* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    auto* o = reinterpret_cast<__nv_bfloat16*>(out);
    switch (d) {
        case 16: linear_attention_kernel<16><<<grid, 256, 0, s>>>(q, mem_kv, o, n, heads, n_mem); return 0;
        case 32: launch_linattn32_tc(qkv, mem_kv, k_shift, out, B, n, heads, n_mem, s); return 0;   // tensor-core path (the only one that takes k_shift)
        case 64: linear_attention_kernel<64><<<grid, 256, 0, s>>>(q, mem_kv, o, n, heads, n_mem); return 0;
        default: return -3;
    }
}

int launch_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const float* mem_k,
                     const float* mem_v, int n_mem, void* out, int B, int nq, int nk, int heads, int d, cudaStream_t s) {
    int threads = 32;
    while (threads < 128 && threads < nq) threads <<= 1;
    const dim3 grid((nq + threads - 1) / threads, heads, B);
    const auto* qq = reinterpret_cast<const __nv_bfloat16*>(q);
    const auto* kk = reinterpret_cast<const __nv_bfloat16*>(k);
    const auto* vv = reinterpret_cast<const __nv_bfloat16*>(v);
    auto* o = reinterpret_cast<__nv_bfloat16*>(out);
    switch (d) {
        case 16: attention_kernel<16><<<grid, threads, 0, s>>>(qq, ldq, kk, ldk, vv, ldv, mem_k, mem_v, n_mem, o, nq, nk, heads); return 0;
        case 32: attention_kernel<32><<<grid, threads, 0, s>>>(qq, ldq, kk, ldk, vv, ldv, mem_k, mem_v, n_mem, o, nq, nk, heads); return 0;
        case 64: attention_kernel<64><<<grid, threads, 0, s>>>(qq, ldq, kk, ldk, vv, ldv, mem_k, mem_v, n_mem, o, nq, nk, heads); return 0;
        default: return -3;
    }
}

}  // namespace ddm
