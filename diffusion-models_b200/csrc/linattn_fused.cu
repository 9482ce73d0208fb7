// Whole LinearAttention block of the U-Net as ONE kernel on tcgen05 / TMEM / TMA (sm_100a).
//
//   y = RMSNorm( W_out . LinAttn( W_qkv . RMSNorm(x) ) + b ) * g  +  x        (denoising_diffusion.py:173-193, :368)
//
// x, y: bf16 [B, n, C] (channels-last pixels of one image are n consecutive rows).  heads = 4, dim_head = 32.
// HBM traffic is one read of x and one write of y (x is read a second time in pass 2, from L2); the 384-channel qkv
// tensor, the attention output and the to_out result never leave the SM.
//
// One CTA per image (persistent over images): 16 epilogue warps + 1 control warp.  The control warp issues every TMA
// load / store and every MMA and hears from the epilogue warps through mbarriers; MMAs for the NEXT step are issued
// before the epilogue of the current one finishes (two accumulator buffers), so the tensor pipe runs under the math.
//
//   pass 1 (64-token chunks):  [K^T | V^T] = W_kv . x_chunk^T          M = 128 (h,d)/(h,e) rows, N = 64 tokens  (TMEM)
//        epilogue: thread = channel row; P = exp2(k * rn[tok] * log2e - shift[c]) and V * rn[tok] as bf16 rows in
//        shared memory.  The softmax over the n tokens (dd:185) needs no running maximum: |k[c][tok]| <= ||w_c|| because
//        the pre-normalised token has unit length, so a per-channel constant shift (host-computed bound) is exact
//        softmax algebra and cannot overflow.
//        ctx[(h,d)][(h',e)] += P_chunk . V_chunk^T                       M = 128, N = 144, K = 64 tokens
//        (all head pairs; only the diagonal blocks are used; column 128 multiplies a row of ones = sum of P)
//   between: ctx + learned memory tokens (dd:181-182), / sum, * d^-0.5  -> block-diagonal bf16 [128][128];
//        M[(h,d)][c] = ctx[(h,d)][(h,e)] . W_out[c][(h,e)]^T  folds to_out into the context (dd:191-192 + conv)
//   pass 2 (128-token tiles):  Q = x_tile . W_q^T  (M = 128 tokens, N = 128); epilogue: thread = token, softmax over
//        the 32 channels of each head (dd:184) -> bf16 rows;  Y = softmax(Q) . M  (N = C);  epilogue: + bias, RMSNorm
//        over C, * g, + x (residual), written in place over the x tile and TMA-stored.
#include "kernels.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {
namespace {

constexpr int kHid = 128;        // heads * dim_head
constexpr int kTileTok = 128;    // tokens per x tile
constexpr int kChunkTok = 64;    // tokens per pass-1 chunk
constexpr int kCtxN = 144;       // context accumulator columns (128 + the ones row padded to a multiple of 16)
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = kEpiThreads + 32;     // + the control warp
// named barriers: 0 = __syncthreads, 1 = epilogue warps only, the rest = epilogue -> control hand-overs
constexpr int kBarEpi = 1, kBarEdone = 2, kBarYdone = 4, kBarCdone = 6, kBarMtdone = 7;

struct LaParams {
    const float* bias_out;   // [C]
    const float* g_out;      // [C]  g * sqrt(C)
    const float* mem_kv;     // [2][4][32][n_mem]
    const float* k_shift;    // [128] per-channel softmax shift (>= max k)
    int B, n, n_mem;
    int trace;               // debugging (env DDM_LAF_TRACE): CTA 0 records (event, clock64) pairs, see scripts/laf_trace.py
};

constexpr int kLafTraceCap = 4096;
__device__ long long g_laf_trace[2 * kLafTraceCap * 2];      // [role: 0 control, 1 epilogue warp 0][event][tag, clock]

struct alignas(8) LaBars {
    uint64_t xfull[4];
    uint64_t accfull[2];     // pass 1: [K^T|V^T] chunk ready; pass 2: Q tile ready
    uint64_t pvdone[2];      // pass 1: context MMA finished reading P/V buffer
    uint64_t yfull[2];       // pass 2: Y tile ready
    uint64_t wfull, woutfull, mdone;
    uint32_t tmem_base;
};

template <int C>
struct LaSmem {
    static constexpr int kAtoms = C / 64;
    static constexpr int kNbuf = 4;
    static constexpr int kXTile = kAtoms * kTileTok * 128;
    static constexpr int kWBytes = kAtoms * 384 * 128;
    static constexpr int off_w = 0;                                   // [atom][384 rows (q|k|v)][128 B]
    static constexpr int off_x = off_w + kWBytes;                     // kNbuf x [atom][128 rows][128 B]
    static constexpr int off_u = off_x + kNbuf * kXTile;              // union region, see below
    static constexpr int kPBytes = 128 * 128;                         // P  [128 ch][64 tok]
    static constexpr int kVBytes = kCtxN * 128;                       // V^T[144 rows][64 tok]
    static constexpr int kUBytes = 2 * kPBytes + 2 * kVBytes;         // 69632
    //   pass 1 : P[0] P[1] V[0] V[1]
    //   between: ctx (bf16 [2 atoms][128][128 B]) at 0
    //   pass 2 : Qs[0] at 0, Qs[1] at 32768 ([2 atoms][128 tok][128 B] each)
    static constexpr int off_mt = off_u + kUBytes;                    // M^T [2 atoms][C rows][128 B]; W_out before the M GEMM
    static constexpr int kMtBytes = 2 * C * 128;
    static constexpr int off_small = off_mt + kMtBytes;
    static constexpr int off_rn = off_small;                          // [kNbuf][128] f32
    static constexpr int off_rnl = off_rn + kNbuf * 128 * 4;          // [kNbuf][128] f32  rn * log2(e)
    static constexpr int off_bias = off_rnl + kNbuf * 128 * 4;        // [C]
    static constexpr int off_g = off_bias + C * 4;                    // [C]
    static constexpr int off_pm = off_g + C * 4;                      // [128][4] exp(mem_k - shift)
    static constexpr int off_mv = off_pm + 128 * 16;                  // [128][4] mem_v
    static constexpr int off_red = off_mv + 128 * 16;                 // [2 tiles][4 parts][128]
    static constexpr int off_bars = off_red + 2 * 4 * 128 * 4;
    static constexpr int kTotal = off_bars + static_cast<int>(sizeof(LaBars));
    static_assert(kTotal + 1024 <= 227 * 1024, "shared-memory plan does not fit");
    static_assert(kUBytes >= 65536, "union region must hold two Qs buffers");
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
template <int C>
__global__ void __launch_bounds__(kThreads, 1)
linattn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmWqkv, const __grid_constant__ CUtensorMap tmWout,
                     const __grid_constant__ LaParams p) {
    using L = LaSmem<C>;
    constexpr int kAtoms = L::kAtoms;
    constexpr int NB = L::kNbuf;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sb = smem_u32(smem);
    LaBars* bars = reinterpret_cast<LaBars*>(smem + L::off_bars);
    float* rn_s = reinterpret_cast<float*>(smem + L::off_rn);
    float* rnl_s = reinterpret_cast<float*>(smem + L::off_rnl);
    float* bias_s = reinterpret_cast<float*>(smem + L::off_bias);
    float* g_s = reinterpret_cast<float*>(smem + L::off_g);
    float* pm_s = reinterpret_cast<float*>(smem + L::off_pm);
    float* mv_s = reinterpret_cast<float*>(smem + L::off_mv);
    float* red_s = reinterpret_cast<float*>(smem + L::off_red);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int q = warp & 3;            // TMEM lane quarter of this warp
    const int part = (warp >> 2) & 3;  // column quarter handled by this (epilogue) warp
    const int row = q * 32 + lane;     // accumulator row of this thread (channel in pass 1, token in pass 2)
    const int sw = row & 7;

    const int T = p.n / kTileTok, J = p.n / kChunkTok;
    const int n_img = (p.B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total_tiles = n_img * 2 * T;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars->xfull[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->accfull[i], 1); mbar_init(&bars->pvdone[i], 1); mbar_init(&bars->yfull[i], 1); }
        mbar_init(&bars->wfull, 1);
        mbar_init(&bars->woutfull, 1);
        mbar_init(&bars->mdone, 1);
        fence_barrier_init();
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmY);
        prefetch_tmap(&tmWqkv);
        prefetch_tmap(&tmWout);
    }
    if (warp == 0) {
        tmem_alloc(&bars->tmem_base, 512u);
        tmem_relinquish();
    }
    for (int i = tid; i < C; i += kThreads) { bias_s[i] = __ldg(p.bias_out + i); g_s[i] = __ldg(p.g_out + i); }
    if (tid < kHid) {        // learned memory tokens (dd:163,181-182): mem_kv [2][h][d][n_mem]
        const float shift = __ldg(p.k_shift + tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool in = j < p.n_mem;
            pm_s[tid * 4 + j] = in ? __expf(__ldg(p.mem_kv + tid * p.n_mem + j) - shift) : 0.0f;
            mv_s[tid * 4 + j] = in ? __ldg(p.mem_kv + (kHid + tid) * p.n_mem + j) : 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float mcl = __ldg(p.k_shift + (row & 127)) * kLog2e;       // pass 1: this thread's channel

    const bool tr_on = p.trace != 0 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kEpiWarps);
    int tr_n = 0;
    auto TR = [&](int ev, int idx) {
        if (tr_on && tr_n < kLafTraceCap) {
            long long* dst = g_laf_trace + (static_cast<size_t>(warp == 0 ? 1 : 0) * kLafTraceCap + tr_n) * 2;
            dst[0] = (static_cast<long long>(ev) << 32) | static_cast<unsigned>(idx);
            dst[1] = clock64();
            ++tr_n;
        }
    };
    auto wait_leader = [&](uint64_t* bar, uint32_t parity) {    // one polling lane per warp
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
    };
    const uint64_t desc0 = umma_desc_sw128(0);
    auto desc = [&](uint32_t addr) -> uint64_t { return desc0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };

    if (warp == kEpiWarps) {
        // =============================================================================================== control warp
        // Issues every TMA load / store and every MMA; never touches data.  It learns that the epilogue warps finished
        // a step through the *done mbarriers (one arrival per epilogue warp) and tells them through the TMA / commit
        // mbarriers, so the epilogue warps never wait for this warp's instruction issue, only for real completions.
        const uint32_t idesc_kv = umma_idesc_bf16(128, kChunkTok);
        const uint32_t idesc_ctx = umma_idesc_bf16(128, kCtxN);
        const uint32_t idesc_128 = umma_idesc_bf16(128, 128);
        const uint32_t idesc_y = umma_idesc_bf16(128, C);
        // x-tile stream: global tile index g = image_iter * 2T + pass * T + t lives in buffer g % NB
        int next_load = 0, released = 0;
        auto pump = [&]() {
            while (next_load < total_tiles && next_load < released + NB) {
                const int g = next_load++;
                if (elect_one()) {
                    const int it = g / (2 * T);
                    const int s = g - it * 2 * T;
                    const int t = s >= T ? s - T : s;
                    const int b = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
                    const int slot = g % NB;
                    mbar_arrive_expect_tx(&bars->xfull[slot], static_cast<uint32_t>(L::kXTile));
                    for (int a = 0; a < kAtoms; ++a)
                        tma_load_2d(smem + L::off_x + slot * L::kXTile + a * (kTileTok * 128), &tmX, &bars->xfull[slot], a * 64,
                                    b * p.n + t * kTileTok);
                }
                __syncwarp();
            }
        };
        auto wait_x = [&](int g) { wait_leader(&bars->xfull[g % NB], static_cast<uint32_t>(g / NB) & 1u); };
        auto issue_kv = [&](int g_tile, int j) {          // chunk j of the image -> accumulator buffer j & 1
            if (elect_one()) {
                const uint32_t xb = sb + L::off_x + (g_tile % NB) * L::kXTile + (j & 1) * (kChunkTok * 128);
                const uint32_t d0 = tmem_base + static_cast<uint32_t>((j & 1) * 128);
#pragma unroll
                for (int kind = 0; kind < 2; ++kind) {
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        const uint64_t ad = desc(sb + L::off_w + a * (384 * 128) + (1 + kind) * (128 * 128));
                        const uint64_t bd = desc(xb + a * (kTileTok * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d0 + kind * 64, ad + 2u * k, bd + 2u * k, idesc_kv, (a | k) ? 1u : 0u);
                    }
                }
                umma_commit(&bars->accfull[j & 1]);
            }
            __syncwarp();
        };
        auto issue_ctx = [&](int j) {
            if (elect_one()) {
                const uint64_t ad = desc(sb + L::off_u + (j & 1) * L::kPBytes);
                const uint64_t bd = desc(sb + L::off_u + 2 * L::kPBytes + (j & 1) * L::kVBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 256u, ad + 2u * k, bd + 2u * k, idesc_ctx, (j | k) ? 1u : 0u);
                umma_commit(&bars->pvdone[j & 1]);
            }
            __syncwarp();
        };
        auto issue_q = [&](int g_tile, int t) {
            if (elect_one()) {
                const uint32_t xb = sb + L::off_x + (g_tile % NB) * L::kXTile;
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    const uint64_t ad = desc(xb + a * (kTileTok * 128));
                    const uint64_t bd = desc(sb + L::off_w + a * (384 * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + static_cast<uint32_t>((t & 1) * 128), ad + 2u * k, bd + 2u * k, idesc_128, (a | k) ? 1u : 0u);
                }
                umma_commit(&bars->accfull[t & 1]);
            }
            __syncwarp();
        };
        auto issue_y = [&](int t) {
            if (elect_one()) {
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint64_t ad = desc(sb + L::off_u + (t & 1) * 32768 + a * 16384);
                    const uint64_t bd = desc(sb + L::off_mt + a * (C * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + 256u + static_cast<uint32_t>((t & 1) * 128), ad + 2u * k, bd + 2u * k, idesc_y, (a | k) ? 1u : 0u);
                }
                umma_commit(&bars->yfull[t & 1]);
            }
            __syncwarp();
        };
        // Epilogue -> control hand-overs are named barriers (the 512 epilogue threads bar.arrive, this warp bar.syncs):
        // every mbarrier poller slows the CTA's other mbarrier traffic, and with 16 + 1 polling warps a wait on an already
        // completed phase took ~700 cycles.  Two ids per event, alternating with the buffer, because the epilogue warps may
        // arrive for step i + 1 before this warp has consumed step i (never for i + 2: that needs an MMA issued after it).
        uint32_t ph_misc = 0;
        auto wait_edone = [&](int i) { named_bar_sync(kBarEdone + i, kThreads); tc_fence_after(); };
        auto wait_ydone = [&](int i) { named_bar_sync(kBarYdone + i, kThreads); tc_fence_after(); };

        // W_out [C][128] is loaded into the M^T region for every image (16 KB from L2, issued as soon as the previous image's
        // last Y MMA has read M^T): the M GEMM reads it before the epilogue warps overwrite the region with M^T
        auto load_wout = [&]() {
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars->woutfull, static_cast<uint32_t>(L::kMtBytes));
                for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_mt + a * (C * 128), &tmWout, &bars->woutfull, a * 64, 0);
            }
            __syncwarp();
        };
        if (elect_one()) {     // W_qkv (pre-norm gain folded in) stays resident
            mbar_arrive_expect_tx(&bars->wfull, static_cast<uint32_t>(L::kWBytes));
            for (int a = 0; a < kAtoms; ++a)
                for (int rb = 0; rb < 3; ++rb)
                    tma_load_2d(smem + L::off_w + a * (384 * 128) + rb * (128 * 128), &tmWqkv, &bars->wfull, a * 64, rb * 128);
        }
        __syncwarp();
        load_wout();
        pump();
        wait_leader(&bars->wfull, 0u);

        int G = 0;
        int stores_in_flight = 0;      // bulk groups committed and not yet waited for
        for (int it = 0; it < n_img; ++it, G += 2 * T) {
            const int b = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            // ---- pass 1
            wait_x(G);
            tc_fence_after();
            issue_kv(G, 0);
            TR(10, 0);
            for (int j = 0; j < J; ++j) {
                if (j + 1 < J) {
                    const int gt = G + ((j + 1) >> 1);
                    if (((j + 1) & 1) == 0) wait_x(gt);
                    TR(9, j + 1);
                    issue_kv(gt, j + 1);        // its accumulator was released by epilogue j - 1 (edone waited below)
                    TR(10, j + 1);
                }
                wait_edone(j & 1);              // P/V of chunk j written, accumulator j & 1 read out
                TR(11, j);
                issue_ctx(j);
                TR(12, j);
                if (j & 1) {                    // both chunks of tile j >> 1 multiplied (epilogue j saw accfull) and its norms taken
                    if (stores_in_flight) { if (elect_one()) bulk_wait_group_read<0>(); __syncwarp(); stores_in_flight = 0; }
                    released = G + (j >> 1) + 1;
                    pump();
                }
            }
            // ---- between the passes: M[(h,d)][c] = ctx[(h,d)][(h',e)] . W_out[c][(h',e)]^T   (W_out sits in the M^T region)
            named_bar_sync(kBarCdone, kThreads);              // block-diagonal context written (every context MMA retired)
            TR(30, it);
            wait_leader(&bars->woutfull, ph_misc & 1u);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint64_t ad = desc(sb + L::off_u + a * 16384);
                    const uint64_t bd = desc(sb + L::off_mt + a * (C * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2u * k, bd + 2u * k, idesc_y, (a | k) ? 1u : 0u);
                }
                umma_commit(&bars->mdone);
            }
            __syncwarp();
            TR(31, it);
            named_bar_sync(kBarMtdone, kThreads);             // M^T written, its accumulator read out
            TR(32, it);
            ph_misc ^= 1u;
            tc_fence_after();
            // ---- pass 2
            const int G2 = G + T;
            wait_x(G2);
            issue_q(G2, 0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) {
                    wait_x(G2 + t + 1);
                    issue_q(G2 + t + 1, t + 1);     // its accumulator was released by epilogue Q t - 1
                }
                TR(20, t);
                wait_edone(t & 1);                  // softmax(q) tile written
                TR(21, t);
                issue_y(t);                         // its accumulator was released by epilogue Y t - 2 (ydone waited below)
                TR(22, t);
                if (t >= 1) {
                    wait_ydone((t - 1) & 1);
                    TR(23, t - 1);
                    if (elect_one()) {
                        // the previous store has read its tile by now: hand that buffer back to the loader first
                        if (stores_in_flight) bulk_wait_group_read<0>();
                        for (int a = 0; a < kAtoms; ++a)
                            tma_store_2d(&tmY, smem + L::off_x + ((G2 + t - 1) % NB) * L::kXTile + a * (kTileTok * 128), a * 64,
                                         b * p.n + (t - 1) * kTileTok);
                        bulk_commit_group();
                    }
                    __syncwarp();
                    if (stores_in_flight) { released = G2 + t - 1; pump(); }
                    stores_in_flight = 1;
                }
            }
            wait_ydone((T - 1) & 1);
            if (elect_one()) {
                if (stores_in_flight) bulk_wait_group_read<0>();
                for (int a = 0; a < kAtoms; ++a)
                    tma_store_2d(&tmY, smem + L::off_x + ((G2 + T - 1) % NB) * L::kXTile + a * (kTileTok * 128), a * 64,
                                 b * p.n + (T - 1) * kTileTok);
                bulk_commit_group();
            }
            __syncwarp();
            if (stores_in_flight) { released = G2 + T - 1; pump(); }
            stores_in_flight = 1;
            if (it + 1 < n_img) load_wout();
        }
        if (elect_one()) bulk_wait_group<0>();
        __syncwarp();
    } else {
        // =============================================================================================== epilogue warps
        // 16 warps: q = warp & 3 is the TMEM lane quarter (accumulator rows 32q .. 32q+31), part = warp >> 2 the column
        // quarter.  Four warps per scheduler: the math of one warp hides the TMEM / shared-memory / MUFU latencies of the
        // others (with 8 warps the kernel was latency-bound at one instruction per 7.7 cycles per warp).
        const float* rn_f = rn_s;
        const float* rnl_f = rnl_s;
        const bool leader = warp == 0 && lane == 0;       // the only mbarrier poller among the epilogue warps
        auto arrive = [&](int bar_id) {     // this thread's generic-proxy writes / TMEM reads are done: tell the control warp
            fence_proxy_async();
            tc_fence_before();
            named_bar_arrive(bar_id, kThreads);
        };
        auto xbar = [&](int g) -> uint64_t* { return &bars->xfull[g % NB]; };
        auto xpar = [&](int g) -> uint32_t { return static_cast<uint32_t>(g / NB) & 1u; };
        // per-token 1 / max(||x||, 1e-12) of tile g (the block's pre-norm, dd:176; gain is folded into W_qkv):
        // four lanes per token, two 16-byte units each
        auto rn_compute = [&](int g) {
            const int slot = g % NB;
            const int r = warp * 8 + (lane >> 2), sub = lane & 3;
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int a = 0; a < kAtoms; ++a) {
                const uint32_t base = sb + L::off_x + slot * L::kXTile + a * (kTileTok * 128) + r * 128;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint4 v = lds_128u(base + static_cast<uint32_t>(((2 * sub + k) ^ (r & 7)) << 4));      // any order: it is a sum
                    s0 = fmaf(bf16_lo(v.x), bf16_lo(v.x), s0); s1 = fmaf(bf16_hi(v.x), bf16_hi(v.x), s1);
                    s0 = fmaf(bf16_lo(v.y), bf16_lo(v.y), s0); s1 = fmaf(bf16_hi(v.y), bf16_hi(v.y), s1);
                    s0 = fmaf(bf16_lo(v.z), bf16_lo(v.z), s0); s1 = fmaf(bf16_hi(v.z), bf16_hi(v.z), s1);
                    s0 = fmaf(bf16_lo(v.w), bf16_lo(v.w), s0); s1 = fmaf(bf16_hi(v.w), bf16_hi(v.w), s1);
                }
            }
            float ss = s0 + s1;
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            if (sub == 0) {
                const float rn = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                rn_s[slot * 128 + r] = rn;
                rnl_s[slot * 128 + r] = rn * kLog2e;
            }
        };
        uint32_t ph_acc = 0, ph_pv = 0, ph_y = 0, ph_m = 0;     // phase bits, one per barrier (tracked by every thread, used by the leader)
        auto poll_acc = [&](int i) { if (leader) mbar_wait(&bars->accfull[i], (ph_acc >> i) & 1u); ph_acc ^= 1u << i; };
        auto poll_pv = [&](int i) { if (leader) mbar_wait(&bars->pvdone[i], (ph_pv >> i) & 1u); ph_pv ^= 1u << i; };
        auto poll_y = [&](int i) { if (leader) mbar_wait(&bars->yfull[i], (ph_y >> i) & 1u); ph_y ^= 1u << i; };
        auto poll_x = [&](int g) { if (leader) mbar_wait(xbar(g), xpar(g)); };
        auto publish = [&]() { named_bar_sync(kBarEpi, kEpiThreads); tc_fence_after(); };     // what the leader saw holds for all

        int G = 0;       // global tile index of the current image's first pass-1 tile
        for (int it = 0; it < n_img; ++it, G += 2 * T) {
            // ======================================================================================= pass 1
            if (tid < 256) {   // the V^T buffers' extra rows: row 128 = ones (its context column is the sum of P), rows 129..143 = 0
                const int vb = tid >> 7, r = (tid & 127) >> 3, u = tid & 7;
                const uint32_t one2 = r == 0 ? 0x3F803F80u : 0u;
                sts_128u(sb + L::off_u + 2 * L::kPBytes + vb * L::kVBytes + (128 + r) * 128 + u * 16, one2, one2, one2, one2);
                // published to the context MMA by the fence + arrival that follows chunk 0 (edone)
            }
            poll_x(G);
            publish();
            rn_compute(G);            // published by the barrier that opens chunk 0
            for (int j = 0; j < J; ++j) {
                const int gt = G + (j >> 1);
                if ((j & 1) && j + 1 < J) poll_x(gt + 1);      // the next tile, for its norms at the end of this chunk
                poll_acc(j & 1);
                if (j >= 2) poll_pv(j & 1);
                publish();
                TR(4, j);
                {   // epilogue: thread = channel `row`; tokens part*16 .. +15 of the chunk
                    const int slot = gt % NB;
                    const uint32_t accb = t_lane + static_cast<uint32_t>((j & 1) * 128 + part * 16);
                    uint32_t kr[16], vr[16];
                    tmem_ld16(accb, kr);
                    tmem_ld16(accb + 64u, vr);
                    const int tok0 = slot * 128 + (j & 1) * 64 + part * 16;
                    float al[16], bl[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a4 = *reinterpret_cast<const float4*>(rnl_f + tok0 + 4 * i);
                        const float4 b4 = *reinterpret_cast<const float4*>(rn_f + tok0 + 4 * i);
                        al[4 * i] = a4.x; al[4 * i + 1] = a4.y; al[4 * i + 2] = a4.z; al[4 * i + 3] = a4.w;
                        bl[4 * i] = b4.x; bl[4 * i + 1] = b4.y; bl[4 * i + 2] = b4.z; bl[4 * i + 3] = b4.w;
                    }
                    tmem_ld_wait();
                    const uint32_t prow = sb + L::off_u + (j & 1) * L::kPBytes + row * 128;
                    const uint32_t vrow = sb + L::off_u + 2 * L::kPBytes + (j & 1) * L::kVBytes + row * 128;
                    uint32_t pw[8], vw[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(kr[2 * i]), al[2 * i], -mcl));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(kr[2 * i + 1]), al[2 * i + 1], -mcl));
                        pw[i] = pack_bf16x2(p0, p1);
                        vw[i] = pack_bf16x2(__uint_as_float(vr[2 * i]) * bl[2 * i], __uint_as_float(vr[2 * i + 1]) * bl[2 * i + 1]);
                    }
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const uint32_t uo = static_cast<uint32_t>(((part * 2 + g2) ^ sw) << 4);
                        sts_128u(prow + uo, pw[4 * g2], pw[4 * g2 + 1], pw[4 * g2 + 2], pw[4 * g2 + 3]);
                        sts_128u(vrow + uo, vw[4 * g2], vw[4 * g2 + 1], vw[4 * g2 + 2], vw[4 * g2 + 3]);
                    }
                }
                TR(5, j);
                arrive(kBarEdone + (j & 1));
                if ((j & 1) && j + 1 < J) rn_compute(gt + 1);
                TR(6, j);
            }
            poll_pv(J & 1);            // chunk J-2
            poll_pv((J - 1) & 1);      // chunk J-1: the context is complete
            publish();
            TR(50, it);

            // ======================================================================================= between the passes
            {   // row (h = q, d = lane) of the block-diagonal bf16 context [128][128]: 16 units of 8 columns, this thread writes
                // units part, part+4, part+8, part+12; exactly one of them (own head, e = 8*part ..) carries data:
                // (ctx + memory tokens) / sum * d^-0.5   (dd:181-182, 187)
                uint32_t cr[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                             : "=r"(cr[0]), "=r"(cr[1]), "=r"(cr[2]), "=r"(cr[3]), "=r"(cr[4]), "=r"(cr[5]), "=r"(cr[6]), "=r"(cr[7])
                             : "r"(t_lane + 256u + static_cast<uint32_t>(q * 32 + part * 8))
                             : "memory");
                float ksum = __uint_as_float(tmem_ld1(t_lane + 256u + 128u));
                tmem_ld_wait();
                const float4 pm = *reinterpret_cast<const float4*>(pm_s + row * 4);
                ksum += (pm.x + pm.y) + (pm.z + pm.w);
                const float sc = 0.17677669529663687f / ksum;     // 32^-0.5 / sum
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[2];
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const int e = part * 8 + 2 * i + e2;
                        const float4 mv = *reinterpret_cast<const float4*>(mv_s + (q * 32 + e) * 4);     // warp-uniform address
                        float c = __uint_as_float(cr[2 * i + e2]);
                        c = fmaf(pm.x, mv.x, c); c = fmaf(pm.y, mv.y, c); c = fmaf(pm.z, mv.z, c); c = fmaf(pm.w, mv.w, c);
                        v[e2] = c * sc;
                    }
                    w[i] = pack_bf16x2(v[0], v[1]);
                }
                const int own = (q >> 1) * 8 + (q & 1) * 4 + part;       // unit index (atom * 8 + unit) of the data
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int U = part + 4 * k;
                    const uint32_t addr = sb + L::off_u + (U >> 3) * 16384 + row * 128 + static_cast<uint32_t>(((U & 7) ^ sw) << 4);
                    if (U == own) sts_128u(addr, w[0], w[1], w[2], w[3]);
                    else sts_128u(addr, 0u, 0u, 0u, 0u);
                }
            }
            arrive(kBarCdone);
            TR(51, it);
            if (leader) mbar_wait(&bars->mdone, ph_m);
            ph_m ^= 1u;
            publish();
            TR(52, it);
            {   // thread = row (h,d) of M; columns c = part * C/4 .. ; stored transposed as M^T[c][(h,d)] (K-major B operand of Y)
                constexpr int kMc = C / 4;
                static_assert(kMc == 16, "M epilogue is written for C = 64");
                const uint32_t mbase = sb + L::off_mt + (row >> 6) * (C * 128) + static_cast<uint32_t>((row & 7) * 2);
                const int ku = (row & 63) >> 3;
                uint32_t mr[16];
                tmem_ld16(t_lane + static_cast<uint32_t>(part * kMc), mr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = part * kMc + i;
                    const unsigned short hv = __bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(mr[i])));
                    asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(mbase + c * 128 + static_cast<uint32_t>((ku ^ (c & 7)) << 4)), "h"(hv) : "memory");
                }
            }
            arrive(kBarMtdone);
            TR(53, it);

            // ======================================================================================= pass 2
            const int G2 = G + T;
            auto epilogue_y = [&](int u_t) {     // its yfull was polled and published by the caller
                TR(45, u_t);
                const int slot = (G2 + u_t) % NB;
                constexpr int kCols = C / 4;                      // columns of this thread: part * kCols .. +kCols-1
                static_assert(kCols == 16, "Y epilogue is written for C = 64");
                uint32_t yr[16];
                tmem_ld16(t_lane + 256u + static_cast<uint32_t>((u_t & 1) * 128 + part * kCols), yr);
                // residual x (own row of the tile); the result is written in place.  column c: atom c / 64, unit (c % 64) / 8
                const uint32_t xrow = sb + L::off_x + slot * L::kXTile + row * 128;
                const uint32_t a0 = xrow + static_cast<uint32_t>(((part * 2) ^ sw) << 4), a1 = xrow + static_cast<uint32_t>(((part * 2 + 1) ^ sw) << 4);
                const uint4 x0 = lds_128u(a0), x1 = lds_128u(a1);
                const float4* bp = reinterpret_cast<const float4*>(bias_s + part * kCols);
                const float4* gp = reinterpret_cast<const float4*>(g_s + part * kCols);
                float v[16];
                tmem_ld_wait();
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bb = bp[i];
                    v[4 * i] = __uint_as_float(yr[4 * i]) + bb.x;
                    v[4 * i + 1] = __uint_as_float(yr[4 * i + 1]) + bb.y;
                    v[4 * i + 2] = __uint_as_float(yr[4 * i + 2]) + bb.z;
                    v[4 * i + 3] = __uint_as_float(yr[4 * i + 3]) + bb.w;
                    s0 = fmaf(v[4 * i], v[4 * i], s0); s1 = fmaf(v[4 * i + 1], v[4 * i + 1], s1);
                    s0 = fmaf(v[4 * i + 2], v[4 * i + 2], s0); s1 = fmaf(v[4 * i + 3], v[4 * i + 3], s1);
                }
                float* red = red_s + (u_t & 1) * 512;
                red[part * 128 + row] = s0 + s1;
                tc_fence_before();
                named_bar_sync(kBarEpi, kEpiThreads);
                TR(46, u_t);
                const float rinv = 1.0f / fmaxf(sqrtf((red[row] + red[128 + row]) + (red[256 + row] + red[384 + row])), 1e-12f);
                const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
                sts_128u(a0, pack_bf16x2(fmaf(v[0] * rinv, g0.x, bf16_lo(x0.x)), fmaf(v[1] * rinv, g0.y, bf16_hi(x0.x))),
                         pack_bf16x2(fmaf(v[2] * rinv, g0.z, bf16_lo(x0.y)), fmaf(v[3] * rinv, g0.w, bf16_hi(x0.y))),
                         pack_bf16x2(fmaf(v[4] * rinv, g1.x, bf16_lo(x0.z)), fmaf(v[5] * rinv, g1.y, bf16_hi(x0.z))),
                         pack_bf16x2(fmaf(v[6] * rinv, g1.z, bf16_lo(x0.w)), fmaf(v[7] * rinv, g1.w, bf16_hi(x0.w))));
                sts_128u(a1, pack_bf16x2(fmaf(v[8] * rinv, g2.x, bf16_lo(x1.x)), fmaf(v[9] * rinv, g2.y, bf16_hi(x1.x))),
                         pack_bf16x2(fmaf(v[10] * rinv, g2.z, bf16_lo(x1.y)), fmaf(v[11] * rinv, g2.w, bf16_hi(x1.y))),
                         pack_bf16x2(fmaf(v[12] * rinv, g3.x, bf16_lo(x1.z)), fmaf(v[13] * rinv, g3.y, bf16_hi(x1.z))),
                         pack_bf16x2(fmaf(v[14] * rinv, g3.z, bf16_lo(x1.w)), fmaf(v[15] * rinv, g3.w, bf16_hi(x1.w))));
                arrive(kBarYdone + (u_t & 1));
                TR(47, u_t);
            };

            poll_x(G2);
            publish();
            rn_compute(G2);           // published by the barrier that opens tile 0
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) poll_x(G2 + t + 1);
                poll_acc(t & 1);
                if (t >= 1) poll_y((t - 1) & 1);
                publish();
                TR(42, t);
                {   // epilogue: thread = token `row`, head `part`: softmax over its 32 channels (dd:184)
                    const int slot = (G2 + t) % NB;
                    const float rnl = rnl_f[slot * 128 + row];
                    uint32_t qr[32];
                    tmem_ld32(t_lane + static_cast<uint32_t>((t & 1) * 128 + part * 32), qr);
                    tmem_ld_wait();
                    float m0 = fmaxf(__uint_as_float(qr[0]), __uint_as_float(qr[1])), m1 = fmaxf(__uint_as_float(qr[2]), __uint_as_float(qr[3]));
                    float m2 = fmaxf(__uint_as_float(qr[4]), __uint_as_float(qr[5])), m3 = fmaxf(__uint_as_float(qr[6]), __uint_as_float(qr[7]));
#pragma unroll
                    for (int i = 8; i < 32; i += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(qr[i]), __uint_as_float(qr[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(qr[i + 2]), __uint_as_float(qr[i + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(qr[i + 4]), __uint_as_float(qr[i + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(qr[i + 6]), __uint_as_float(qr[i + 7])));
                    }
                    const float mm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * rnl;
                    float e[32];
                    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        e[i] = ex2_approx(fmaf(__uint_as_float(qr[i]), rnl, -mm));
                        e[i + 1] = ex2_approx(fmaf(__uint_as_float(qr[i + 1]), rnl, -mm));
                        e[i + 2] = ex2_approx(fmaf(__uint_as_float(qr[i + 2]), rnl, -mm));
                        e[i + 3] = ex2_approx(fmaf(__uint_as_float(qr[i + 3]), rnl, -mm));
                        s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3];
                    }
                    const float inv = 1.0f / ((s0 + s1) + (s2 + s3));
                    const uint32_t qrow = sb + L::off_u + (t & 1) * 32768 + (part >> 1) * 16384 + row * 128;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        sts_128u(qrow + static_cast<uint32_t>((((part & 1) * 4 + u) ^ sw) << 4),
                                 pack_bf16x2(e[8 * u] * inv, e[8 * u + 1] * inv), pack_bf16x2(e[8 * u + 2] * inv, e[8 * u + 3] * inv),
                                 pack_bf16x2(e[8 * u + 4] * inv, e[8 * u + 5] * inv), pack_bf16x2(e[8 * u + 6] * inv, e[8 * u + 7] * inv));
                }
                TR(43, t);
                arrive(kBarEdone + (t & 1));
                TR(44, t);
                if (t + 1 < T) rn_compute(G2 + t + 1);      // published by epilogue_y's barrier / the next tile's opening barrier
                if (t >= 1) epilogue_y(t - 1);
            }
            poll_y((T - 1) & 1);
            publish();
            epilogue_y(T - 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512u);
    }
}

}  // namespace

int linattn_fused_prepare_attributes() {
    return static_cast<int>(cudaFuncSetAttribute(linattn_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 LaSmem<64>::kTotal + 1024));
}

int linattn_trace_read(long long* host, int cap) {
    static long long tmp[2 * kLafTraceCap * 2];
    cudaMemcpyFromSymbol(tmp, g_laf_trace, sizeof(tmp));
    int n = 0;
    for (int i = 0; i < 2 * kLafTraceCap && n < cap; ++i)
        if (tmp[2 * i + 1] != 0) { host[3 * n] = i / kLafTraceCap; host[3 * n + 1] = tmp[2 * i]; host[3 * n + 2] = tmp[2 * i + 1]; ++n; }
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, g_laf_trace);
    cudaMemset(sym, 0, sizeof(tmp));
    return n;
}

bool linattn_fused_supported(int C, int n, int heads, int d, int n_mem) {
    return C == 64 && heads == 4 && d == 32 && n >= kTileTok && (n % kTileTok) == 0 && n_mem >= 0 && n_mem <= 4;
}

void launch_linattn_fused(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmWqkv, const CUtensorMap& tmWout,
                          const float* bias_out, const float* g_out, const float* mem_kv, const float* k_shift, int B, int n, int C,
                          int n_mem, int num_sms, int trace, cudaStream_t s) {
    LaParams p;
    p.bias_out = bias_out; p.g_out = g_out; p.mem_kv = mem_kv; p.k_shift = k_shift;
    p.B = B; p.n = n; p.n_mem = n_mem;
    p.trace = trace;
    const int grid = B < num_sms ? B : num_sms;
    linattn_fused_kernel<64><<<grid, kThreads, LaSmem<64>::kTotal + 1024, s>>>(tmX, tmY, tmWqkv, tmWout, p);
}

}  // namespace ddm
