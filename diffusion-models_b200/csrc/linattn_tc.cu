// Linear attention core for dim_head = 32 on the (legacy) mma.sync tensor path.
//   LinearAttention.forward, denoising_diffusion.py:178-192:
//     q = softmax_d(q) * d^-0.5 ; k = softmax_n(k) over the n tokens + 4 learned memory tokens ;
//     context[d][e] = sum_n k[d][n] v[e][n] ; out[e][n] = sum_d context[d][e] q[d][n]
// The op is <2 % of the network FLOPs and bound by reading the 384-channel qkv tensor once (k twice), so the two tiny
// GEMMs (32x32xn and nx32x32) run as bf16 mma.sync.m16n8k16 fed by ldmatrix -- tcgen05 would need 128-row tiles and
// TMEM for 32x32 outputs.  One CTA per (batch, head), 4 warps, 128-token tiles; softmax statistics in fp32.
#include "kernels.cuh"

#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {
namespace {

constexpr int D = 32;        // dim_head
constexpr int TOK = 128;     // tokens per tile == threads per CTA
constexpr int PITCH = 40;    // bf16 elements per smem row (80 B: conflict-free ldmatrix)

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// exp(x - m) as ex2.approx(x * log2e - m * log2e): one FFMA + one MUFU (expf() expands to ~6 instructions with its
// range handling; the exponentials are ~60 % of this kernel's instruction stream)
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void load_row32(const __nv_bfloat16* p, float (&f)[32]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 u = __ldg(q + c);
        f[8 * c + 0] = bf16_lo(u.x); f[8 * c + 1] = bf16_hi(u.x); f[8 * c + 2] = bf16_lo(u.y); f[8 * c + 3] = bf16_hi(u.y);
        f[8 * c + 4] = bf16_lo(u.z); f[8 * c + 5] = bf16_hi(u.z); f[8 * c + 6] = bf16_lo(u.w); f[8 * c + 7] = bf16_hi(u.w);
    }
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
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

__global__ void __launch_bounds__(TOK)
linattn32_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ mem_kv, const float* __restrict__ k_shift,
                    __nv_bfloat16* __restrict__ out, int n, int heads, int n_mem) {
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int HD = heads * D, ld = 3 * HD;
    const __nv_bfloat16* base = qkv + static_cast<long long>(b) * n * ld;
    const __nv_bfloat16* qp = base + h * D;
    const __nv_bfloat16* kp = base + HD + h * D;
    const __nv_bfloat16* vp = base + 2 * HD + h * D;
    const float* mk = mem_kv + (static_cast<long long>(0) * heads + h) * D * n_mem;   // [d][n_mem]
    const float* mv = mem_kv + (static_cast<long long>(1) * heads + h) * D * n_mem;

    __shared__ __align__(16) __nv_bfloat16 PV[2 * TOK * PITCH];
    __nv_bfloat16* Ps = PV;                                    // exp(k - max) tile, later softmax(q) tile
    __nv_bfloat16* Vs = PV + TOK * PITCH;                      // v tile, later the output tile
    __shared__ __align__(16) __nv_bfloat16 Cs[D * PITCH];     // normalised context [d][e], q scale folded in
    __shared__ float red[TOK * 33];                           // reductions / per-warp partial contexts
    __shared__ float kmax[D], ksum[D];
    float* scratch2 = reinterpret_cast<float*>(PV);            // [TOK][33] floats overlaying the P/V tiles
    static_assert(sizeof(__nv_bfloat16) * 2 * TOK * PITCH >= sizeof(float) * TOK * 33, "scratch overlay too small");

    // Global access pattern: a token's 32 channels of one head are 64 contiguous bytes; thread t handles 16-byte
    // part (t & 3) of token (t >> 2) + 32 i of a tile, so that a warp request touches 8 cache lines instead of 32
    // (one thread per token made the kernel L1-wavefront-bound: l1tex 90 % busy, profiles/r01_ncu_linattn*.txt).
    const int part = tid & 3;            // channels 8*part .. 8*part+7
    const int trow = tid >> 2;           // token within a group of 32

    // ---- pass 1: per-channel max of k over the tokens (softmax over n, dd:185) -- or, when the caller supplies an upper bound of
    // |k| per channel (pre-normalised tokens are unit vectors: |k[c]| <= ||w_c||, packing.linattn_k_shift), that bound as the
    // shift: softmax is invariant to it, and one of the three latency-exposed passes over the tokens disappears
    if (k_shift != nullptr) {
        if (tid < D) kmax[tid] = __ldg(k_shift + h * D + tid);
        __syncthreads();
    } else {
        float mx[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = -INFINITY;
        // eight independent 16-byte loads in flight per thread (the pass is pure latency otherwise)
        for (int tok = trow; tok < n; tok += 32 * 8) {
            uint4 u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (tok + 32 * i < n) u[i] = __ldg(reinterpret_cast<const uint4*>(kp + static_cast<long long>(tok + 32 * i) * ld + part * 8));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (tok + 32 * i < n) {
                    float f[8];
                    unpack8(u[i], f);
#pragma unroll
                    for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], f[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) red[trow * 33 + part * 8 + c] = mx[c];
        __syncthreads();
        if (tid < D) {
            float m = -INFINITY;
            for (int t = 0; t < 32; ++t) m = fmaxf(m, red[t * 33 + tid]);
            for (int j = 0; j < n_mem; ++j) m = fmaxf(m, __ldg(mk + tid * n_mem + j));
            kmax[tid] = m;
        }
        __syncthreads();
    }

    // ---- pass 2: context[d][e] = sum_tok exp(k[tok][d] - max[d]) * v[tok][e]     (dd:189)
    float cacc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) cacc[i][j][e] = 0.0f;
    float psum[8];
    float km[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { psum[c] = 0.0f; km[c] = kmax[part * 8 + c] * kLog2e; }     // pre-scaled by log2(e)

    // the next tile's k/v rows are fetched into registers while the current tile is multiplied (global latency
    // was the limiter: 16 warps/SM, long-scoreboard stalls 8 per issue)
    uint4 kreg[TOK / 32], vreg[TOK / 32];
    auto fetch_kv = [&](int t0) {
#pragma unroll
        for (int i = 0; i < TOK / 32; ++i) {
            const int tok = t0 + trow + 32 * i;
            if (tok < n) {
                kreg[i] = __ldg(reinterpret_cast<const uint4*>(kp + static_cast<long long>(tok) * ld + part * 8));
                vreg[i] = __ldg(reinterpret_cast<const uint4*>(vp + static_cast<long long>(tok) * ld + part * 8));
            }
        }
    };
    fetch_kv(0);
    for (int t0 = 0; t0 < n; t0 += TOK) {
#pragma unroll
        for (int i = 0; i < TOK / 32; ++i) {
            const int lt = trow + 32 * i;               // token inside the tile
            const int tok = t0 + lt;
            uint4* pdst = reinterpret_cast<uint4*>(Ps + lt * PITCH + part * 8);
            uint4* vdst = reinterpret_cast<uint4*>(Vs + lt * PITCH + part * 8);
            if (tok < n) {
                float f[8];
                unpack8(kreg[i], f);
                uint32_t w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    w[c] = pack_bf16x2(ex2_approx(fmaf(f[2 * c], kLog2e, -km[2 * c])), ex2_approx(fmaf(f[2 * c + 1], kLog2e, -km[2 * c + 1])));
                    psum[2 * c] += bf16_lo(w[c]);          // normalise with exactly the rounded weights the MMA sees
                    psum[2 * c + 1] += bf16_hi(w[c]);
                }
                *pdst = make_uint4(w[0], w[1], w[2], w[3]);
                *vdst = vreg[i];
            } else {
                *pdst = make_uint4(0, 0, 0, 0);
                *vdst = make_uint4(0, 0, 0, 0);
            }
        }
        if (t0 + TOK < n) fetch_kv(t0 + TOK);
        __syncthreads();
        // warp w contracts its 32 tokens: A = P^T (stored [tok][d] -> ldmatrix.trans), B = V (stored [tok][e] -> .trans)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int k0 = warp * 32 + ks * 16;
            const int i = lane >> 3, j = lane & 7;
            uint32_t a[2][4], bfr[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldmatrix_x4_trans(a[mt], Ps + (k0 + j + 8 * (i >> 1)) * PITCH + mt * 16 + 8 * (i & 1));
#pragma unroll
            for (int np = 0; np < 2; ++np)       // two n-tiles per ldmatrix.x4
                ldmatrix_x4_trans(bfr[np], Vs + (k0 + j + 8 * (i & 1)) * PITCH + np * 16 + 8 * (i >> 1));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16(cacc[mt][nt], a[mt], bfr[nt >> 1][2 * (nt & 1)], bfr[nt >> 1][2 * (nt & 1) + 1]);
        }
        __syncthreads();
    }
    // combine: per-warp partial contexts + per-thread partial sums + the learned memory tokens (dd:181-182)
    {
        const int g = lane >> 2, t = lane & 3;
        float* wpart = red + warp * (D * D);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int d0 = mt * 16 + g, e0 = nt * 8 + 2 * t;
                wpart[d0 * D + e0] = cacc[mt][nt][0];
                wpart[d0 * D + e0 + 1] = cacc[mt][nt][1];
                wpart[(d0 + 8) * D + e0] = cacc[mt][nt][2];
                wpart[(d0 + 8) * D + e0 + 1] = cacc[mt][nt][3];
            }
#pragma unroll
        for (int c = 0; c < 8; ++c) scratch2[trow * 33 + part * 8 + c] = psum[c];
        __syncthreads();
        if (tid < D) {
            float s = 0.0f;
            for (int r = 0; r < 32; ++r) s += scratch2[r * 33 + tid];
            for (int j = 0; j < n_mem; ++j) s += __expf(__ldg(mk + tid * n_mem + j) - kmax[tid]);
            ksum[tid] = s;
        }
        __syncthreads();
        const float qscale = rsqrtf(static_cast<float>(D));       // dd:187, folded into the context
        const int d = tid >> 2, eb = (tid & 3) * 8;
        float cv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float s = red[d * D + eb + e] + red[D * D + d * D + eb + e] + red[2 * D * D + d * D + eb + e] + red[3 * D * D + d * D + eb + e];
            for (int j = 0; j < n_mem; ++j)
                s = fmaf(__expf(__ldg(mk + d * n_mem + j) - kmax[d]), __ldg(mv + (eb + e) * n_mem + j), s);
            cv[e] = s * qscale / ksum[d];
        }
        __syncthreads();      // scratch2 (overlaying Ps/Vs) is dead from here on
        *reinterpret_cast<uint4*>(Cs + d * PITCH + eb) =
            make_uint4(pack_bf16x2(cv[0], cv[1]), pack_bf16x2(cv[2], cv[3]), pack_bf16x2(cv[4], cv[5]), pack_bf16x2(cv[6], cv[7]));
        __syncthreads();
    }

    // ---- pass 3: out[tok][e] = sum_d softmax_d(q)[tok][d] * context[d][e]          (dd:184,191)
    uint32_t cb[2][2][4];      // context B fragments: [k-step][n-pair][regs], loaded once
    {
        const int i = lane >> 3, j = lane & 7;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldmatrix_x4_trans(cb[ks][np], Cs + (ks * 16 + j + 8 * (i & 1)) * PITCH + np * 16 + 8 * (i >> 1));
    }
    uint4 qreg[TOK / 32];
    auto fetch_q = [&](int t0) {
#pragma unroll
        for (int i = 0; i < TOK / 32; ++i) {
            const int tok = t0 + trow + 32 * i;
            if (tok < n) qreg[i] = __ldg(reinterpret_cast<const uint4*>(qp + static_cast<long long>(tok) * ld + part * 8));
        }
    };
    fetch_q(0);
    for (int t0 = 0; t0 < n; t0 += TOK) {
#pragma unroll
        for (int i = 0; i < TOK / 32; ++i) {
            const int lt = trow + 32 * i;
            const int tok = t0 + lt;
            float f[8];
            float m = -INFINITY;
            if (tok < n) {
                unpack8(qreg[i], f);
#pragma unroll
                for (int c = 0; c < 8; ++c) m = fmaxf(m, f[c]);
            }
            // softmax over the 32 channels of the token = the 4 lanes that share it (dd:184)
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float ssum = 0.0f;
            if (tok < n) {
                const float ml = m * kLog2e;
#pragma unroll
                for (int c = 0; c < 8; ++c) { f[c] = ex2_approx(fmaf(f[c], kLog2e, -ml)); ssum += f[c]; }
            }
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
            uint4* qdst = reinterpret_cast<uint4*>(Ps + lt * PITCH + part * 8);
            if (tok < n) {
                const float inv = 1.0f / ssum;
                *qdst = make_uint4(pack_bf16x2(f[0] * inv, f[1] * inv), pack_bf16x2(f[2] * inv, f[3] * inv),
                                   pack_bf16x2(f[4] * inv, f[5] * inv), pack_bf16x2(f[6] * inv, f[7] * inv));
            } else {
                *qdst = make_uint4(0, 0, 0, 0);
            }
        }
        if (t0 + TOK < n) fetch_q(t0 + TOK);
        __syncthreads();
        float oacc[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) oacc[i][j][e] = 0.0f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int m0 = warp * 32 + mt * 16;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                ldmatrix_x4(a, Ps + (m0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + ks * 16 + 8 * (lane >> 4));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16(oacc[mt][nt], a, cb[ks][nt >> 1][2 * (nt & 1)], cb[ks][nt >> 1][2 * (nt & 1) + 1]);
            }
        }
        {
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int r0 = warp * 32 + mt * 16 + g, e0 = nt * 8 + 2 * t;
                    *reinterpret_cast<uint32_t*>(Vs + r0 * PITCH + e0) = pack_bf16x2(oacc[mt][nt][0], oacc[mt][nt][1]);
                    *reinterpret_cast<uint32_t*>(Vs + (r0 + 8) * PITCH + e0) = pack_bf16x2(oacc[mt][nt][2], oacc[mt][nt][3]);
                }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TOK / 32; ++i) {
            const int lt = trow + 32 * i;
            const int tok = t0 + lt;
            if (tok < n)
                *reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * n + tok) * HD + h * D + part * 8) =
                    *reinterpret_cast<const uint4*>(Vs + lt * PITCH + part * 8);
        }
        __syncthreads();
    }
}

}  // namespace

void launch_linattn32_tc(const void* qkv, const float* mem_kv, const float* k_shift, void* out, int B, int n, int heads, int n_mem,
                         cudaStream_t s) {
    const dim3 grid(heads, B);
    linattn32_tc_kernel<<<grid, TOK, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), mem_kv, k_shift,
                                             reinterpret_cast<__nv_bfloat16*>(out), n, heads, n_mem);
}

}  // namespace ddm
