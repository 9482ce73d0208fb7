// Softmax attention on tcgen05 / TMEM / TMA (sm_100a): S = Q K^T in TMEM, softmax in registers, O = P V in TMEM.
//
// Replaces the op chain of Attention + Attend (denoising_diffusion.py:220-228, attend.py:109-124), of CrossAttention's core
// (denoising_diffusion_text_conditional.py:66-77) and of the VAE decoder's AttnBlock (ldm/modules/diffusionmodules/model.py:
// 190-215):  out = softmax(q k^T d^-0.5) v  per (image, head), with optional learned memory keys/values prepended.
//
// q: bf16 rows [B*nq] (row stride ldq, head h at column h*d), k / v likewise over [B*nk] rows, out: bf16 [B*nq][heads*d].
// A CTA owns 128 consecutive QUERY ROWS of the flattened matrix and one head.  When an image has fewer than 128 queries
// (16 at the 32-px bottleneck) the tile simply spans several images and the keys of all of them: a query only attends the
// keys of its own image, which is a column-range mask per row -- the tensor core multiplies the full 128 x 128 tile either
// way, and no 128-row tile is wasted on 16 rows.  Keys are walked in tiles of 128; the softmax is two-pass (pass 0: running
// max / sum per row from S alone; pass 1: S again, P = exp(s - max) / sum as a bf16 A operand, O += P V), so the O
// accumulator in TMEM never needs rescaling.  V is consumed as it lies in memory ([key][d], d contiguous) as an MN-major
// B operand: no transposed copy.
#include "kernels.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {
namespace {

constexpr int kQT = 128;          // query rows per CTA
constexpr int kKT = 128;          // keys per tile
constexpr int kMemPad = 16;       // learned memory keys padded to one K = 16 MMA step

struct AtParams {
    const float* mem_k;           // [heads][n_mem][d] or null
    const float* mem_v;
    __nv_bfloat16* out;
    int B, nq, nk, heads, n_mem;
    float scale_log2;             // d^-0.5 * log2(e)
};

template <int D>
struct AtSmem {
    static constexpr int kAtoms = (D + 63) / 64;
    static constexpr int off_q = 0;                                  // [atom][128 rows][128 B]
    static constexpr int off_k = off_q + kAtoms * 16384;             // [atom][128 keys][128 B]          K-major B of S
    static constexpr int off_v = off_k + kAtoms * 16384;             // [d atom][128 keys][128 B]        MN-major B of O
    static constexpr int off_p = off_v + kAtoms * 16384;             // [2 atoms][128 rows][128 B]       K-major A of O (keys = K)
    static constexpr int off_km = off_p + 2 * 16384;                 // [atom][16 rows][128 B]   memory keys
    static constexpr int off_vm = off_km + kAtoms * 2048;            // [d atom][16 keys][128 B] memory values
    static constexpr int off_pm = off_vm + kAtoms * 2048;            // [128 rows][128 B]        P of the memory keys (first 32 B)
    static constexpr int off_bars = off_pm + 16384;
    static constexpr int kTotal = off_bars + 64;
    static constexpr int kTmemCols = D > 64 ? 512 : 256;
    static constexpr int kOCol = D > 64 ? 256 : 192;                 // O accumulator columns; S at 0..127, memory S at 128..143
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// MN-major operand (rows = K index, 64 MN elements = 128 B contiguous per row), 128-byte swizzle: 8-row groups along K are
// 1024 B apart (SBO); the leading-dimension offset (next 64 MN elements) is not used with N <= 64.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// bounded wait: a protocol bug traps within ~0.5 s (cudaErrorLaunchFailure) instead of spinning for minutes
__device__ __forceinline__ void wait_tagged(uint64_t* bar, uint32_t parity, int /*site*/) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 1000000000ll) __trap();
    }
}

template <int D>
__global__ void __launch_bounds__(128)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AtParams p) {
    using L = AtSmem<D>;
    constexpr int kAtoms = L::kAtoms;
    constexpr int kDN = D < 64 ? D : 64;              // N of one O MMA (one 64-column block of d)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sb = smem_u32(smem);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + L::off_bars);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y;
    const long long r0 = static_cast<long long>(blockIdx.x) * kQT;
    const long long rows_q = static_cast<long long>(p.B) * p.nq, rows_k = static_cast<long long>(p.B) * p.nk;
    const int sw = tid & 7;

    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmV);
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, static_cast<uint32_t>(L::kTmemCols));
        tmem_relinquish();
    }
    // learned memory keys / values (dd:223-224) as bf16 operand tiles; rows >= n_mem are zero
    const bool has_mem = p.n_mem > 0;
    for (int i = tid; i < kMemPad * (D / 8); i += 128) {
        const int r = i / (D / 8), c8 = i - r * (D / 8);       // 8 consecutive channels of memory row r
        uint32_t kw[4] = {0u, 0u, 0u, 0u}, vw[4] = {0u, 0u, 0u, 0u};
        if (has_mem && r < p.n_mem) {
            const float* mk = p.mem_k + (static_cast<long long>(h) * p.n_mem + r) * D + c8 * 8;
            const float* mv = p.mem_v + (static_cast<long long>(h) * p.n_mem + r) * D + c8 * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) { kw[j] = pack2(__ldg(mk + 2 * j), __ldg(mk + 2 * j + 1)); vw[j] = pack2(__ldg(mv + 2 * j), __ldg(mv + 2 * j + 1)); }
        }
        const int a = c8 >> 3, u = c8 & 7;
        sts_128u(sb + L::off_km + a * 2048 + r * 128 + static_cast<uint32_t>((u ^ (r & 7)) << 4), kw[0], kw[1], kw[2], kw[3]);
        sts_128u(sb + L::off_vm + a * 2048 + r * 128 + static_cast<uint32_t>((u ^ (r & 7)) << 4), vw[0], vw[1], vw[2], vw[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

    // this thread's query row and the key rows it may attend (its own image)
    const long long r = r0 + tid;
    const bool row_ok = r < rows_q;
    const long long bq = row_ok ? r / p.nq : 0;
    const long long kmin = bq * p.nk, kmax = kmin + p.nk;
    // key rows any row of this tile may attend
    const long long b_lo = r0 / p.nq;
    long long b_hi = (r0 + kQT - 1) / p.nq;
    if (b_hi > p.B - 1) b_hi = p.B - 1;
    const long long k_begin = b_lo * p.nk, k_end = (b_hi + 1) * p.nk;
    const int n_tiles = static_cast<int>((k_end - k_begin + kKT - 1) / kKT);

    const uint64_t dK = umma_desc_sw128(0), dMN = umma_desc_sw128_mn(0);
    auto kdesc = [&](uint32_t addr) -> uint64_t { return dK | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
    auto mndesc = [&](uint32_t addr) -> uint64_t { return dMN | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
    const uint32_t idesc_s = umma_idesc_bf16(128, kKT);
    const uint32_t idesc_sm = umma_idesc_bf16(128, kMemPad);
    const uint32_t idesc_o = umma_idesc_bf16(128, kDN) | (1u << 16);      // B operand MN-major

    uint32_t ph_tma = 0, ph_mma = 0;
    if (tid == 0) {     // Q tile (zero-filled past the last row / last column)
        mbar_arrive_expect_tx(bar_tma, static_cast<uint32_t>(kAtoms * 16384));
        for (int a = 0; a < kAtoms; ++a) tma_load_2d(smem + L::off_q + a * 16384, &tmQ, bar_tma, h * D + a * 64, static_cast<int>(r0));
    }
    wait_tagged(bar_tma, ph_tma, 1);
    ph_tma ^= 1u;

    float m_run = -INFINITY, l_run = 0.0f, inv_l = 0.0f;
    bool o_started = false;
    for (int pass = 0; pass < 2; ++pass) {
        for (int kt = 0; kt < n_tiles; ++kt) {
            const long long k0 = k_begin + static_cast<long long>(kt) * kKT;
            const bool with_mem = has_mem && kt == 0;
            __syncthreads();                // the previous tile's operands / accumulator columns are free
            if (tid == 0) {
                tc_fence_after();
                mbar_arrive_expect_tx(bar_tma, static_cast<uint32_t>((pass ? 2 : 1) * kAtoms * 16384));
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_2d(smem + L::off_k + a * 16384, &tmK, bar_tma, h * D + a * 64, static_cast<int>(k0));
                    if (pass) tma_load_2d(smem + L::off_v + a * 16384, &tmV, bar_tma, h * D + a * 64, static_cast<int>(k0));
                }
                wait_tagged(bar_tma, ph_tma, 2);
                // S = Q K^T (+ the memory keys' columns)
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t off = static_cast<uint32_t>((ks >> 2) * 16384 + (ks & 3) * 32);
                    umma_bf16(tmem_base, kdesc(sb + L::off_q + off), kdesc(sb + L::off_k + off), idesc_s, ks ? 1u : 0u);
                }
                if (with_mem) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t off = static_cast<uint32_t>((ks >> 2) * 16384 + (ks & 3) * 32);
                        const uint32_t offm = static_cast<uint32_t>((ks >> 2) * 2048 + (ks & 3) * 32);
                        umma_bf16(tmem_base + 128u, kdesc(sb + L::off_q + off), kdesc(sb + L::off_km + offm), idesc_sm, ks ? 1u : 0u);
                    }
                }
                umma_commit(bar_mma);
            }
            ph_tma ^= 1u;
            wait_tagged(bar_mma, ph_mma, 3);
            ph_mma ^= 1u;
            tc_fence_after();

            // valid key columns of this tile for this row
            long long lo_l = kmin - k0, hi_l = kmax - k0;
            const int lo = row_ok ? static_cast<int>(lo_l < 0 ? 0 : (lo_l > kKT ? kKT : lo_l)) : 0;
            const int hi = row_ok ? static_cast<int>(hi_l < 0 ? 0 : (hi_l > kKT ? kKT : hi_l)) : 0;
            // tcgen05.ld is warp-collective: whether a 32-column block is read at all must be decided per WARP (the union of
            // its lanes' ranges; rows of two images, or past the end, share a warp), the per-lane range only masks values
            const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : kKT);
            const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
            if (pass == 0) {
                float tm = -INFINITY;
                float sm[kMemPad];
                if (with_mem) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + 128u, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < kMemPad; ++j) {
                        sm[j] = j < p.n_mem ? __uint_as_float(v[j]) * p.scale_log2 : -INFINITY;
                        tm = fmaxf(tm, sm[j]);
                    }
                }
                // two sweeps over the tile's columns: max, then sum (S stays in TMEM; nothing is kept in registers)
                for (int c0 = 0; c0 < kKT; c0 += 32) {
                    if (c0 + 32 <= wlo || c0 >= whi) continue;
                    uint32_t v[32];
                    tmem_ld32(t_lane + static_cast<uint32_t>(c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j >= lo && c0 + j < hi) tm = fmaxf(tm, __uint_as_float(v[j]) * p.scale_log2);
                }
                const float m_new = fmaxf(m_run, tm);
                const bool any = m_new > -INFINITY;           // false only for a row with no key at all so far (e.g. past the end)
                const float m_use = any ? m_new : 0.0f;
                float add = 0.0f;
                if (with_mem) {
#pragma unroll
                    for (int j = 0; j < kMemPad; ++j) add += ex2f(sm[j] - m_use);
                }
                for (int c0 = 0; c0 < kKT; c0 += 32) {
                    if (c0 + 32 <= wlo || c0 >= whi) continue;
                    uint32_t v[32];
                    tmem_ld32(t_lane + static_cast<uint32_t>(c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j >= lo && c0 + j < hi) add += ex2f(fmaf(__uint_as_float(v[j]), p.scale_log2, -m_use));
                }
                if (any) {
                    l_run = l_run * ex2f(m_run - m_new) + add;
                    m_run = m_new;
                }
                tc_fence_before();
            } else {
                if (kt == 0) { inv_l = l_run > 0.0f ? 1.0f / l_run : 0.0f; if (!(m_run > -INFINITY)) m_run = 0.0f; }
                // P = exp(s - max) / sum as bf16, K-major (keys along K): two 64-key atoms of [128 rows][128 B]
                if (with_mem) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + 128u, v);
                    tmem_ld_wait();
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = (2 * j < p.n_mem) ? ex2f(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_run)) * inv_l : 0.0f;
                        const float b2 = (2 * j + 1 < p.n_mem) ? ex2f(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_run)) * inv_l : 0.0f;
                        w[j] = pack2(a, b2);
                    }
                    const uint32_t prow = sb + L::off_pm + tid * 128;
                    sts_128u(prow + static_cast<uint32_t>((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
                    sts_128u(prow + static_cast<uint32_t>((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
                }
                for (int c0 = 0; c0 < kKT; c0 += 32) {
                    uint32_t w[16];
                    if (c0 + 32 <= wlo || c0 >= whi) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) w[j] = 0u;
                    } else {
                        uint32_t v[32];
                        tmem_ld32(t_lane + static_cast<uint32_t>(c0), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int ca = c0 + 2 * j, cb = ca + 1;
                            const float a = (ca >= lo && ca < hi) ? ex2f(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_run)) * inv_l : 0.0f;
                            const float b2 = (cb >= lo && cb < hi) ? ex2f(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_run)) * inv_l : 0.0f;
                            w[j] = pack2(a, b2);
                        }
                    }
                    const uint32_t prow = sb + L::off_p + (c0 >> 6) * 16384 + tid * 128;
                    const int u0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        sts_128u(prow + static_cast<uint32_t>(((u0 + u) ^ sw) << 4), w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {     // O += P V  (+ the memory values), one 64-column block of d per MMA
                    tc_fence_after();
#pragma unroll
                    for (int da = 0; da < kAtoms; ++da) {
                        const uint32_t dcol = tmem_base + static_cast<uint32_t>(L::kOCol + da * 64);
                        uint32_t acc = o_started ? 1u : 0u;
                        if (with_mem) {
                            umma_bf16(dcol, kdesc(sb + L::off_pm), mndesc(sb + L::off_vm + da * 2048), idesc_o, acc);
                            acc = 1u;
                        }
#pragma unroll
                        for (int ks = 0; ks < kKT / 16; ++ks) {
                            const uint32_t aoff = static_cast<uint32_t>((ks >> 2) * 16384 + (ks & 3) * 32);
                            umma_bf16(dcol, kdesc(sb + L::off_p + aoff), mndesc(sb + L::off_v + da * 16384 + ks * 2048), idesc_o, acc);
                            acc = 1u;
                        }
                    }
                    umma_commit(bar_mma);
                }
                o_started = true;
                wait_tagged(bar_mma, ph_mma, 4);     // P / V / S are reused by the next tile
                ph_mma ^= 1u;
                tc_fence_after();
            }
        }
    }
    // O row -> bf16 -> global
    if (n_tiles > 0) {
        __nv_bfloat16* orow = p.out + r * (static_cast<long long>(p.heads) * D) + h * D;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_lane + static_cast<uint32_t>(L::kOCol + c0), v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(orow + c0 + u * 8) =
                        make_uint4(pack2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])), pack2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                                   pack2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])), pack2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(L::kTmemCols));
    }
}

}  // namespace

int attention_tc_prepare_attributes() {
    int r = static_cast<int>(cudaFuncSetAttribute(attention_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem<32>::kTotal + 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem<64>::kTotal + 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(attention_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem<128>::kTotal + 1024));
    return r;
}

bool attention_tc_supported(int d, int n_mem) { return (d == 32 || d == 64 || d == 128) && n_mem >= 0 && n_mem <= kMemPad; }

void launch_attention_tc(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const float* mem_k, const float* mem_v,
                         int n_mem, void* out, int B, int nq, int nk, int heads, int d, cudaStream_t s) {
    AtParams p;
    p.mem_k = mem_k; p.mem_v = mem_v; p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.B = B; p.nq = nq; p.nk = nk; p.heads = heads; p.n_mem = n_mem;
    p.scale_log2 = rsqrtf(static_cast<float>(d)) * 1.4426950408889634f;
    const long long rows = static_cast<long long>(B) * nq;
    const dim3 grid(static_cast<unsigned>((rows + kQT - 1) / kQT), static_cast<unsigned>(heads));
    switch (d) {
        case 32: attention_tc_kernel<32><<<grid, 128, AtSmem<32>::kTotal + 1024, s>>>(tmQ, tmK, tmV, p); break;
        case 64: attention_tc_kernel<64><<<grid, 128, AtSmem<64>::kTotal + 1024, s>>>(tmQ, tmK, tmV, p); break;
        default: attention_tc_kernel<128><<<grid, 128, AtSmem<128>::kTotal + 1024, s>>>(tmQ, tmK, tmV, p); break;
    }
}

}  // namespace ddm
