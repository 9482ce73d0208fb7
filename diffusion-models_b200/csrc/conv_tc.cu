// See conv_tc.cuh for the design.  sm_100a only: tcgen05.mma / TMEM / TMA.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace ddm {

namespace {

struct alignas(8) ConvBarriers {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kATileBytes + p.block_n * (kChunkK * 2);
    ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(smem + p.num_stages * stage_bytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], 128);
        }
        fence_barrier_init();
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmW);
    }
    if (warp == 1) {
        tmem_alloc(&bars->tmem_base, static_cast<uint32_t>(p.tmem_cols));
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const int chunks_per_tap = p.chunks0 + p.chunks1;
    const int k_chunks = p.ntaps * chunks_per_tap;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_tile = tile / p.m_tiles;
                const int m_tile = tile - n_tile * p.m_tiles;
                const int tx = m_tile % p.tiles_x;
                const int ty = (m_tile / p.tiles_x) % p.tiles_y;
                const int tb = m_tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
                const int n0 = n_tile * p.block_n;
                int kcol = 0;
                for (int t = 0; t < p.ntaps; ++t) {
                    const int cx = x0 + p.tap_dx[t], cy = y0 + p.tap_dy[t], cp = p.tap_p[t];
                    for (int c = 0; c < chunks_per_tap; ++c) {
                        mbar_wait(&bars->empty[stage], phase ^ 1u);
                        uint8_t* a_dst = smem + stage * stage_bytes;
                        uint8_t* b_dst = a_dst + kATileBytes;
                        mbar_arrive_expect_tx(&bars->full[stage], static_cast<uint32_t>(stage_bytes));
                        if (c < p.chunks0) {
                            tma_load_5d(a_dst, &tmA0, &bars->full[stage], c * kChunkK, cx, cp, cy, b0);
                        } else {
                            tma_load_5d(a_dst, &tmA1, &bars->full[stage], (c - p.chunks0) * kChunkK, cx, cp, cy, b0);
                        }
                        tma_load_2d(b_dst, &tmW, &bars->full[stage], kcol, n0);
                        kcol += kChunkK;
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kTileM, static_cast<uint32_t>(p.block_n));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                    const uint64_t a_desc = umma_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_desc_sw128(a_addr + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kChunkK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr>>4) field
                        umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&bars->acc_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;            // accumulator row == tile pixel
        const int bx = r % p.bw;
        const int by = (r / p.bw) % p.bh;
        const int bi = r / (p.bw * p.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int n_tile = tile / p.m_tiles;
            const int m_tile = tile - n_tile * p.m_tiles;
            const int tx = m_tile % p.tiles_x;
            const int ty = (m_tile / p.tiles_x) % p.tiles_y;
            const int tb = m_tile / (p.tiles_x * p.tiles_y);
            const int x = tx * p.bw + bx, y = ty * p.bh + by, b = tb * p.bb + bi;
            const int n0 = n_tile * p.block_n;
            const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
            const int oyy = y * p.sy + p.oy, oxx = x * p.sx + p.ox;
            const long long out_pix = (static_cast<long long>(b) * p.OH + oyy) * p.OW + oxx;
            const float rs = (p.row_scale != nullptr && valid)
                                 ? __ldg(p.row_scale + (static_cast<long long>(b) * p.H + y) * p.W + x) : 1.0f;
            const float* ss = (p.scale_shift != nullptr) ? p.scale_shift + static_cast<long long>(valid ? b : 0) * p.ss_stride
                                                         : nullptr;

            mbar_wait(&bars->acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.acc_stride);
            const int ncols = min(p.block_n, p.N - n0);      // real columns of this N tile (multiple of 16 not required)
            const int nchunks = (ncols + 15) >> 4;

            float rinv = 1.0f;
            if (p.norm_g != nullptr) {
                float sumsq = 0.0f;
                for (int c = 0; c < nchunks; ++c) {
                    uint32_t v[16];
                    tmem_ld16(t_row + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + c * 16 + j;
                        if (n < p.N) {
                            float f = __uint_as_float(v[j]) * rs;
                            if (p.bias != nullptr) f += __ldg(p.bias + n);
                            sumsq = fmaf(f, f, sumsq);
                        }
                    }
                }
                rinv = 1.0f / fmaxf(sqrtf(sumsq), 1e-12f);
            }

            float out_sumsq = 0.0f;
            for (int c = 0; c < nchunks; ++c) {
                uint32_t v[16];
                tmem_ld16(t_row + c * 16, v);
                tmem_ld_wait();
                if (c == nchunks - 1) {
                    // last TMEM read of this accumulator stage: hand it back to the MMA warp before the stores
                    tc_fence_before();
                    mbar_arrive(&bars->acc_empty[acc]);
                }
                float f[16];
                const int nb = n0 + c * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = nb + j;
                    float t = 0.0f;
                    if (n < p.N) {
                        t = __uint_as_float(v[j]) * rs;
                        if (p.bias != nullptr) t += __ldg(p.bias + n);
                        if (p.norm_g != nullptr) t = t * rinv * __ldg(p.norm_g + n);
                        if (ss != nullptr) t = fmaf(t, __ldg(ss + n) + 1.0f, __ldg(ss + p.N + n));
                        if (p.act == 1) t = silu_f(t);
                    }
                    f[j] = t;
                }
                if (!valid) continue;
                if (p.out_f32_nchw) {
                    float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = nb + j;
                        if (n < p.N) o[((static_cast<long long>(b) * p.N + n) * p.OH + oyy) * p.OW + oxx] = f[j];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_pix * p.ld_out + nb;
                    const bool full16 = (nb + 16 <= p.N);
                    if (p.residual != nullptr) {
                        const __nv_bfloat16* rp = p.residual + out_pix * p.ld_res + nb;
                        if (full16) {
                            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp));
                            const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                f[2 * j] += bf16_lo(rr[j]);
                                f[2 * j + 1] += bf16_hi(rr[j]);
                            }
                        } else {
                            for (int j = 0; j < 16 && nb + j < p.N; ++j) f[j] += __bfloat162float(rp[j]);
                        }
                    }
                    if (full16) {
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                        if (p.rnorm_out != nullptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float a = bf16_lo(w[j]), c2 = bf16_hi(w[j]);
                                out_sumsq = fmaf(a, a, fmaf(c2, c2, out_sumsq));
                            }
                        }
                        reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
                        for (int j = 0; j < 16 && nb + j < p.N; ++j) {
                            const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
                            const float a = __bfloat162float(h);
                            out_sumsq = fmaf(a, a, out_sumsq);
                            o[j] = h;
                        }
                    }
                }
            }
            if (p.rnorm_out != nullptr && valid) p.rnorm_out[out_pix] = 1.0f / fmaxf(sqrtf(out_sumsq), 1e-12f);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
}

}  // namespace

int conv_smem_bytes(int block_n, int num_stages) {
    return num_stages * (kATileBytes + block_n * kChunkK * 2) + static_cast<int>(sizeof(ConvBarriers)) + 1024;
}

int conv_prepare_attributes() {
    return static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void launch_conv(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW, const ConvParams& p,
                 int num_sms, cudaStream_t stream) {
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    const int smem = conv_smem_bytes(p.block_n, p.num_stages);
    conv_tc_kernel<<<grid, kConvThreads, smem, stream>>>(tmA0, tmA1, tmW, p);
}

}  // namespace ddm
