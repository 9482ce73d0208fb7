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
constexpr int kThreads = kEpiThreads + 96;     // + the three control warps (TMA, K issuer, Cx issuer)
constexpr int kGroupThreads = kEpiThreads / 2; // two epilogue groups of 8 warps
// named barriers: 0 = __syncthreads, 1 / 2 = the two epilogue groups, 8 = all epilogue warps, the rest = epilogue -> control
// (edone -> both issuers, release / ydone -> the TMA warp, cdone -> Cx, mtdone -> both issuers)
// Over-run safety (a bar.arrive for the next use of an id before the consumer passed the current one is undefined):
//   edone : the group's next arrival needs an MMA that K issues after passing this one, and the barrier only completes once Cx
//           (and in pass 1 the TMA warp) arrived too -> two ids (buffer parity) suffice;
//   ydone : consumed by the TMA warp alone; with a ring of kNbuf = 4 tiles, tile t + 4 can only be loaded after the TMA warp
//           stored tile t, i.e. passed ydone(t) -> four ids (t & 3) can never be over-run.
//   afree : the group's next arrival needs the accumulator that K fills after passing this one -> two ids.
constexpr int kBarEpi = 1, kBarEdone = 3, kBarYdone = 5, kBarCdone = 9, kBarMtdone = 10, kBarAll = 11, kBarAfree = 12;
constexpr int kPass1Edone = kGroupThreads + 64, kPass2Edone = kGroupThreads + 32;    // group + (Cx, TMA) / + Cx
constexpr int kAfree = kGroupThreads + 32;                                            // group + K

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
    uint64_t xfull[8];
    uint64_t accfull[2];     // pass 1: [K^T|V^T] chunk ready; pass 2: Q tile ready
    uint64_t pvdone[2];      // pass 1: context MMA finished reading P/V buffer
    uint64_t yfull[2];       // pass 2: Y tile ready
    uint64_t wfull, woutfull, mdone;
    uint64_t wqfull;         // C = 128: W_q of this image landed (wfull then means W_k | W_v of this image)
    uint32_t tmem_base;
};

template <int C>
struct LaSmem {
    static constexpr int kAtoms = C / 64;
    // C = 64: W_qkv (48 KB) stays resident, W_out is loaded per image into the M^T region, ring of 4 x-tiles.
    // C = 128: the weights do not fit next to the x ring, so the 64 KB weight region is phased per image -- pass 1: W_k | W_v;
    //          pass 2: W_q (first half) and M^T (second half) -- W_out goes to the union region between the passes, and the
    //          ring holds 2 x-tiles of 32 KB.  (The per-image weight reloads come from L2: 0.1 GB per launch at B = 1024.)
    static constexpr bool kResidentW = C == 64;
    static constexpr int kNbuf = kResidentW ? 4 : 2;
    static constexpr int kXTile = kAtoms * kTileTok * 128;
    static constexpr int kWBytes = kResidentW ? kAtoms * 384 * 128 : 2 * kAtoms * 128 * 128;
    static constexpr int off_w = 0;
    static constexpr int off_x = off_w + kWBytes;                     // kNbuf x [atom][128 rows][128 B]
    static constexpr int off_u = off_x + kNbuf * kXTile;              // union region, see below
    static constexpr int kPBytes = 128 * 128;                         // P  [128 ch][64 tok]
    static constexpr int kVBytes = kCtxN * 128;                       // V^T[144 rows][64 tok]
    static constexpr int kUBytes = 2 * kPBytes + 2 * kVBytes;         // 69632
    //   pass 1 : P[0] P[1] V[0] V[1]
    //   between: ctx (bf16 [2 atoms][128][128 B]) at 0  (C = 128: W_out [2 atoms][128 rows][128 B] at 32768)
    //   pass 2 : Qs[0] at 0, Qs[1] at 32768 ([2 atoms][128 tok][128 B] each)
    static constexpr int kMtBytes = 2 * C * 128;                      // M^T [2 atoms][C rows][128 B]
    static constexpr int off_mt = kResidentW ? off_u + kUBytes : off_w + kAtoms * 128 * 128;
    static constexpr int off_wout = kResidentW ? off_mt : off_u + 32768;      // W_out [2 atoms][C rows][128 B] before the M GEMM
    static constexpr int off_small = kResidentW ? off_mt + kMtBytes : off_u + kUBytes;
    static constexpr int off_rn = off_small;                          // [kNbuf][128] f32
    static constexpr int off_rnl = off_rn + kNbuf * 128 * 4;          // [kNbuf][128] f32  rn * log2(e)
    static constexpr int off_bias = off_rnl + kNbuf * 128 * 4;        // [C]
    static constexpr int off_g = off_bias + C * 4;                    // [C]
    static constexpr int off_pm = off_g + C * 4;                      // [128][4] exp(mem_k - shift)
    static constexpr int off_mv = off_pm + 128 * 16;                  // [128][4] mem_v
    static constexpr int off_red = off_mv + 128 * 16;                 // [2 groups][2 halves][128]
    static constexpr int off_bars = off_red + 2 * 2 * 128 * 4;
    static constexpr int kTotal = off_bars + static_cast<int>(sizeof(LaBars));
    static_assert(kTotal + 1024 <= 227 * 1024, "shared-memory plan does not fit");
    static_assert(kUBytes >= 65536, "union region must hold two Qs buffers");
    // shared-memory offset of a 128-row weight block: kind 0 = W_q, 1 = W_k, 2 = W_v; atom = 64-channel slice of the input
    static constexpr __host__ __device__ int w_block(int kind, int a) {
        return kResidentW ? off_w + a * (384 * 128) + kind * (128 * 128)
                          : (kind == 0 ? off_w + a * (128 * 128) : off_w + ((kind - 1) * kAtoms + a) * (128 * 128));
    }
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

    griddep_launch();          // programmatic dependent launch (ptx.cuh): the next kernel may begin its prologue
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars->xfull[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->accfull[i], 1); mbar_init(&bars->pvdone[i], 1); mbar_init(&bars->yfull[i], 1); }
        mbar_init(&bars->wfull, 1);
        mbar_init(&bars->woutfull, 1);
        mbar_init(&bars->mdone, 1);
        mbar_init(&bars->wqfull, 1);
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

    const bool tr_on = p.trace != 0 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kEpiWarps + 1);
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

    if (warp >= kEpiWarps) {
        // =============================================================================================== control warps
        // Three single-purpose warps that never touch data:
        //   T  (warp 16): every TMA load and store -- W_qkv once, W_out per image, the x-tile ring, the y tiles;
        //   K  (warp 17): the MMAs that FILL the per-group accumulators: [K^T|V^T] chunks (pass 1), Q tiles (pass 2);
        //   Cx (warp 18): the MMAs that CONSUME a finished epilogue: context (pass 1), M (between), Y (pass 2).
        // A tcgen05.mma blocks its issuing thread while the ~4-deep MMA queue is full, i.e. for about the execution time of the
        // batch (8 x 48 cycles for a chunk): one issuer for everything was the critical path (1.9 k cycles per chunk for 0.7 k of
        // tensor work).  They hear from the epilogue groups through named barriers (bar.arrive by the group, bar.sync here:
        // every mbarrier poller slows the CTA's mbarrier traffic, a wait on a completed phase took ~700 cycles with 17 pollers)
        // and talk back through the TMA / tcgen05.commit mbarriers.  Barrier ids alternate with the buffer because a group may
        // arrive for step i + 2 only after an MMA that is issued behind step i's bar.sync.
        const uint32_t idesc_kv = umma_idesc_bf16(128, kChunkTok);
        const uint32_t idesc_ctx = umma_idesc_bf16(128, kCtxN);
        const uint32_t idesc_128 = umma_idesc_bf16(128, 128);
        const uint32_t idesc_y = umma_idesc_bf16(128, C);
        auto wait_x = [&](int g) { wait_leader(&bars->xfull[g % NB], static_cast<uint32_t>(g / NB) & 1u); };
        auto wait_edone = [&](int i, int count) { named_bar_sync(kBarEdone + i, count); tc_fence_after(); };

        if (warp == kEpiWarps) {
            // ------------------------------------------------------------------------------------------- T: TMA
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
            // W_out [C][128] goes into the M^T region for every image (16 KB from L2, as soon as the previous image's last Y MMA
            // has read M^T): the M GEMM reads it before the epilogue warps overwrite the region with M^T
            auto load_wout = [&]() {
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->woutfull, static_cast<uint32_t>(L::kMtBytes));
                    for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_wout + a * (C * 128), &tmWout, &bars->woutfull, a * 64, 0);
                }
                __syncwarp();
            };
            // 128-row blocks of W_qkv (pre-norm gain folded in): kinds [k0, k1) of (q, k, v), all input-channel atoms
            auto load_w = [&](uint64_t* bar, int k0, int k1) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar, static_cast<uint32_t>((k1 - k0) * kAtoms * 128 * 128));
                    for (int kind = k0; kind < k1; ++kind)
                        for (int a = 0; a < kAtoms; ++a)
                            tma_load_2d(smem + L::w_block(kind, a), &tmWqkv, bar, a * 64, kind * 128);
                }
                __syncwarp();
            };
            if (L::kResidentW) { load_w(&bars->wfull, 0, 3); load_wout(); }      // resident for the CTA's lifetime / first image
            else load_w(&bars->wfull, 1, 3);                                      // W_k | W_v of the first image
            griddep_wait();        // the weights above are constants; x is the previous kernel's output
            pump();
            int G = 0;
            for (int it = 0; it < n_img; ++it, G += 2 * T) {
                const int b = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
                for (int j = 0; j < J; ++j) {       // pass 1: a tile is free once both of its chunks went through their epilogues
                    named_bar_sync(kBarEdone + (j & 1), kPass1Edone);
                    if (j & 1) { released = G + (j >> 1) + 1; pump(); }
                }
                if (!L::kResidentW) load_w(&bars->wqfull, 0, 1);     // every [K^T|V^T] MMA has retired: W_q over the W_k half
                const int G2 = G + T;
                for (int t = 0; t < T; ++t) {       // pass 2: y tile (written in place over its x tile) -> global
                    named_bar_sync(kBarYdone + (t & 3), kGroupThreads + 32);
                    if (elect_one()) {
                        for (int a = 0; a < kAtoms; ++a)
                            tma_store_2d(&tmY, smem + L::off_x + ((G2 + t) % NB) * L::kXTile + a * (kTileTok * 128), a * 64, b * p.n + t * kTileTok);
                        bulk_commit_group();
                        bulk_wait_group_read<0>();      // the buffer goes back to the loader as soon as the store has read it
                    }
                    __syncwarp();
                    released = G2 + t + 1;
                    pump();
                }
                if (it + 1 < n_img) {               // every Q / Y MMA of this image has retired (their epilogues ran)
                    if (L::kResidentW) load_wout();
                    else load_w(&bars->wfull, 1, 3);
                }
            }
            if (elect_one()) bulk_wait_group<0>();
            __syncwarp();
        } else if (warp == kEpiWarps + 1) {
            // ------------------------------------------------------------------------------------------- K: accumulator-filling MMAs
            auto issue_kv = [&](int g_tile, int j) {          // chunk j of the image -> accumulator buffer j & 1
                if (elect_one()) {
                    const uint32_t xb = sb + L::off_x + (g_tile % NB) * L::kXTile + (j & 1) * (kChunkTok * 128);
                    const uint32_t d0 = tmem_base + static_cast<uint32_t>((j & 1) * 128);
#pragma unroll
                    for (int kind = 0; kind < 2; ++kind) {
#pragma unroll
                        for (int a = 0; a < kAtoms; ++a) {
                            const uint64_t ad = desc(sb + L::w_block(1 + kind, a));
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
            auto issue_q = [&](int g_tile, int t) {
                if (elect_one()) {
                    const uint32_t xb = sb + L::off_x + (g_tile % NB) * L::kXTile;
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        const uint64_t ad = desc(xb + a * (kTileTok * 128));
                        const uint64_t bd = desc(sb + L::w_block(0, a));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + static_cast<uint32_t>((t & 1) * 128), ad + 2u * k, bd + 2u * k, idesc_128, (a | k) ? 1u : 0u);
                    }
                    umma_commit(&bars->accfull[t & 1]);
                }
                __syncwarp();
            };
            if (L::kResidentW) wait_leader(&bars->wfull, 0u);
            int G = 0;
            for (int it = 0; it < n_img; ++it, G += 2 * T) {
                if (!L::kResidentW) wait_leader(&bars->wfull, static_cast<uint32_t>(it) & 1u);      // this image's W_k | W_v
                wait_x(G);
                tc_fence_after();
                issue_kv(G, 0);
                TR(10, 0);
                issue_kv(G, 1);
                for (int j = 0; j < J; ++j) {
                    named_bar_sync(kBarAfree + (j & 1), kAfree);   // accumulator j & 1 read out by its group (mid-epilogue)
                    tc_fence_after();
                    TR(11, j);
                    if (j + 2 < J) {
                        const int gt = G + ((j + 2) >> 1);
                        if ((j & 1) == 0) wait_x(gt);
                        issue_kv(gt, j + 2);
                        TR(10, j + 2);
                    }
                }
                named_bar_sync(kBarMtdone, kEpiThreads + 64);      // M^T written: accumulator columns 0..63 are free again
                tc_fence_after();
                const int G2 = G + T;
                if (!L::kResidentW) wait_leader(&bars->wqfull, static_cast<uint32_t>(it) & 1u);     // this image's W_q
                wait_x(G2);
                issue_q(G2, 0);
                if (T > 1) { wait_x(G2 + 1); issue_q(G2 + 1, 1); }
                for (int t = 0; t < T; ++t) {
                    named_bar_sync(kBarAfree + (t & 1), kAfree);   // Q accumulator t & 1 read out (mid-epilogue)
                    tc_fence_after();
                    TR(21, t);
                    if (t + 2 < T) { wait_x(G2 + t + 2); issue_q(G2 + t + 2, t + 2); }
                }
            }
        } else {
            // ------------------------------------------------------------------------------------------- Cx: epilogue-consuming MMAs
            uint32_t ph_w = 0;
            for (int it = 0; it < n_img; ++it) {
                for (int j = 0; j < J; ++j) {
                    wait_edone(j & 1, kPass1Edone);              // P / V^T of chunk j written
                    if (elect_one()) {
                        const uint64_t ad = desc(sb + L::off_u + (j & 1) * L::kPBytes);
                        const uint64_t bd = desc(sb + L::off_u + 2 * L::kPBytes + (j & 1) * L::kVBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 256u, ad + 2u * k, bd + 2u * k, idesc_ctx, (j | k) ? 1u : 0u);
                        umma_commit(&bars->pvdone[j & 1]);
                    }
                    __syncwarp();
                    TR(12, j);
                }
                // between the passes: M[(h,d)][c] = ctx[(h,d)][(h',e)] . W_out[c][(h',e)]^T   (W_out sits in the M^T region)
                named_bar_sync(kBarCdone, kEpiThreads + 32);      // block-diagonal context written (every context MMA retired)
                TR(30, it);
                if (!L::kResidentW) {       // the V^T buffers are dead now: W_out into the union region (16-32 KB from L2)
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&bars->woutfull, static_cast<uint32_t>(L::kMtBytes));
                        for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_wout + a * (C * 128), &tmWout, &bars->woutfull, a * 64, 0);
                    }
                    __syncwarp();
                }
                wait_leader(&bars->woutfull, ph_w);
                ph_w ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const uint64_t ad = desc(sb + L::off_u + a * 16384);
                        const uint64_t bd = desc(sb + L::off_wout + a * (C * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2u * k, bd + 2u * k, idesc_y, (a | k) ? 1u : 0u);
                    }
                    umma_commit(&bars->mdone);
                }
                __syncwarp();
                TR(31, it);
                named_bar_sync(kBarMtdone, kEpiThreads + 64);     // M^T written, its accumulator read out
                tc_fence_after();
                for (int t = 0; t < T; ++t) {
                    wait_edone(t & 1, kPass2Edone);              // softmax(q) tile t written; Y accumulator t & 1 was read out two tiles ago
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
                    TR(22, t);
                }
            }
        }
    } else {
        // =============================================================================================== epilogue warps
        // Two groups of 8 warps.  Group g owns the pass-1 chunks j with j & 1 == g and the pass-2 tiles t with t & 1 == g:
        // accumulator buffer g, P / V^T buffer g, softmax(q) buffer g and Y accumulator g are its private property, so the
        // two groups never synchronise with each other inside a pass and the fixed latencies of one (mbarrier poll, named
        // barrier, fences, MMA round trip) are covered by the other's math.  Inside a group: q = lane quarter (accumulator
        // rows 32q .. 32q+31), half = column half.  Only the group leader polls mbarriers (every poller slows the CTA's
        // mbarrier traffic); the rest of the group learns through the group's named barrier.
        const int grp = warp >> 3;
        const int wg = warp & 7;
        const int half = wg >> 2;
        const int gtid = tid & (kGroupThreads - 1);
        const bool leader = wg == 0 && lane == 0;
        const int bar_g = kBarEpi + grp;
        const float* rn_f = rn_s;
        const float* rnl_f = rnl_s;
        auto arrive = [&](int bar_id, int count) {     // this thread's generic-proxy writes / TMEM reads are done: tell the control warp
            fence_proxy_async();
            tc_fence_before();
            named_bar_arrive(bar_id, count);
        };
        auto xbar = [&](int g) -> uint64_t* { return &bars->xfull[g % NB]; };
        auto xpar = [&](int g) -> uint32_t { return static_cast<uint32_t>(g / NB) & 1u; };
        auto poll_x = [&](int g) { if (leader) mbar_wait(xbar(g), xpar(g)); };
        auto publish = [&]() { named_bar_sync(bar_g, kGroupThreads); tc_fence_after(); };     // what the leader saw holds for the group
        // per-token 1 / max(||x||, 1e-12) (the block's pre-norm, dd:176; gain is folded into W_qkv) of `ntok` tokens starting at
        // token `tok0` of tile g, by this group: 256 / ntok lanes per token
        auto rn_compute = [&](int g, int tok0, int ntok) {
            const int slot = g % NB;
            const int lpt = kGroupThreads / ntok;                  // 4 (64-token chunk) or 2 (128-token tile)
            const int r = tok0 + gtid / lpt, sub = gtid % lpt;
            const int upl = 8 / lpt;                               // 16-byte units per lane
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int a = 0; a < kAtoms; ++a) {
                const uint32_t base = sb + L::off_x + slot * L::kXTile + a * (kTileTok * 128) + r * 128;
                for (int k = 0; k < upl; ++k) {
                    const uint4 v = lds_128u(base + static_cast<uint32_t>(((upl * sub + k) ^ (r & 7)) << 4));      // any order: it is a sum
                    s0 = fmaf(bf16_lo(v.x), bf16_lo(v.x), s0); s1 = fmaf(bf16_hi(v.x), bf16_hi(v.x), s1);
                    s0 = fmaf(bf16_lo(v.y), bf16_lo(v.y), s0); s1 = fmaf(bf16_hi(v.y), bf16_hi(v.y), s1);
                    s0 = fmaf(bf16_lo(v.z), bf16_lo(v.z), s0); s1 = fmaf(bf16_hi(v.z), bf16_hi(v.z), s1);
                    s0 = fmaf(bf16_lo(v.w), bf16_lo(v.w), s0); s1 = fmaf(bf16_hi(v.w), bf16_hi(v.w), s1);
                }
            }
            float ss = s0 + s1;
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            if (lpt == 4) ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            if (sub == 0) {
                const float rn = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                rn_s[slot * 128 + r] = rn;
                rnl_s[slot * 128 + r] = rn * kLog2e;
            }
        };
        uint32_t ph_acc = 0, ph_pv = 0, ph_y = 0, ph_m = 0;     // phases of THIS group's barriers (accfull[grp], pvdone[grp], yfull[grp])

        int G = 0;       // global tile index of the current image's first pass-1 tile
        for (int it = 0; it < n_img; ++it, G += 2 * T) {
            // ======================================================================================= pass 1
            if (gtid < 128) {   // this group's V^T buffer: row 128 = ones (its context column is the sum of P), rows 129..143 = 0
                const int r = gtid >> 3, u = gtid & 7;
                const uint32_t one2 = r == 0 ? 0x3F803F80u : 0u;
                sts_128u(sb + L::off_u + 2 * L::kPBytes + grp * L::kVBytes + (128 + r) * 128 + u * 16, one2, one2, one2, one2);
                // published to the context MMA by the fence + arrival that follows this group's first chunk (edone)
            }
            poll_x(G);
            publish();
            rn_compute(G, grp * kChunkTok, kChunkTok);            // own first chunk; published by the barrier that opens it
            for (int j = grp; j < J; j += 2) {
                const int gt = G + (j >> 1);
                if (j + 2 < J) poll_x(gt + 1);                    // the tile of this group's next chunk, for its norms
                if (leader) mbar_wait(&bars->accfull[grp], ph_acc);
                ph_acc ^= 1u;
                if (j >= 2) { if (leader) mbar_wait(&bars->pvdone[grp], ph_pv); ph_pv ^= 1u; }
                publish();
                TR(4, j);
                {   // epilogue: thread = channel `row`; tokens half*32 .. +31 of the chunk, in two halves of 16 (register budget)
                    const int slot = gt % NB;
                    const uint32_t prow = sb + L::off_u + grp * L::kPBytes + row * 128;
                    const uint32_t vrow = sb + L::off_u + 2 * L::kPBytes + grp * L::kVBytes + row * 128;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const uint32_t accb = t_lane + static_cast<uint32_t>(grp * 128 + half * 32 + hh * 16);
                        uint32_t kr[16], vr[16];
                        tmem_ld16(accb, kr);
                        tmem_ld16(accb + 64u, vr);
                        const int tok0 = slot * 128 + grp * 64 + half * 32 + hh * 16;
                        float al[16], bl[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 a4 = *reinterpret_cast<const float4*>(rnl_f + tok0 + 4 * i);
                            const float4 b4 = *reinterpret_cast<const float4*>(rn_f + tok0 + 4 * i);
                            al[4 * i] = a4.x; al[4 * i + 1] = a4.y; al[4 * i + 2] = a4.z; al[4 * i + 3] = a4.w;
                            bl[4 * i] = b4.x; bl[4 * i + 1] = b4.y; bl[4 * i + 2] = b4.z; bl[4 * i + 3] = b4.w;
                        }
                        tmem_ld_wait();
                        if (hh == 1) {      // the accumulator is in registers: K may refill it while the rest of the math runs
                            tc_fence_before();
                            named_bar_arrive(kBarAfree + grp, kAfree);
                        }
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
                            const uint32_t uo = static_cast<uint32_t>(((half * 4 + hh * 2 + g2) ^ sw) << 4);
                            sts_128u(prow + uo, pw[4 * g2], pw[4 * g2 + 1], pw[4 * g2 + 2], pw[4 * g2 + 3]);
                            sts_128u(vrow + uo, vw[4 * g2], vw[4 * g2 + 1], vw[4 * g2 + 2], vw[4 * g2 + 3]);
                        }
                    }
                }
                TR(5, j);
                arrive(kBarEdone + grp, kPass1Edone);        // (the tile of this chunk is not read again: its norms were taken before)
                if (j + 2 < J) rn_compute(gt + 1, grp * kChunkTok, kChunkTok);
                TR(6, j);
            }
            // this group's last context MMA (chunk J - 2 + grp), then both groups meet: the context is complete
            if (leader) mbar_wait(&bars->pvdone[grp], ph_pv);
            ph_pv ^= 1u;
            named_bar_sync(kBarAll, kEpiThreads);
            tc_fence_after();
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
            arrive(kBarCdone, kEpiThreads + 32);
            TR(51, it);
            if (warp == 0 && lane == 0) mbar_wait(&bars->mdone, ph_m);
            ph_m ^= 1u;
            named_bar_sync(kBarAll, kEpiThreads);
            tc_fence_after();
            TR(52, it);
            {   // thread = row (h,d) of M; columns c = part * C/4 .. ; stored transposed as M^T[c][(h,d)] (K-major B operand of Y)
                constexpr int kMc = C / 4;
                const uint32_t mbase = sb + L::off_mt + (row >> 6) * (C * 128) + static_cast<uint32_t>((row & 7) * 2);
                const int ku = (row & 63) >> 3;
#pragma unroll
                for (int c16 = 0; c16 < kMc / 16; ++c16) {
                    uint32_t mr[16];
                    tmem_ld16(t_lane + static_cast<uint32_t>(part * kMc + c16 * 16), mr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = part * kMc + c16 * 16 + i;
                        const unsigned short hv = __bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(mr[i])));
                        asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(mbase + c * 128 + static_cast<uint32_t>((ku ^ (c & 7)) << 4)), "h"(hv) : "memory");
                    }
                }
            }
            arrive(kBarMtdone, kEpiThreads + 64);
            TR(53, it);

            // ======================================================================================= pass 2
            const int G2 = G + T;
            if (grp < T) {
                poll_x(G2 + grp);
                publish();
                rn_compute(G2 + grp, 0, kTileTok);      // own first tile; published by the barrier that opens it
            }
            for (int t = grp; t < T; t += 2) {
                const int slot = (G2 + t) % NB;
                if (leader) mbar_wait(&bars->accfull[grp], ph_acc);
                ph_acc ^= 1u;
                publish();
                TR(42, t);
                {   // Q epilogue: thread = token `row`; heads 2*half, 2*half+1: softmax over the 32 channels of each (dd:184)
                    const float rnl = rnl_f[slot * 128 + row];
                    const uint32_t qrow = sb + L::off_u + grp * 32768 + half * 16384 + row * 128;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t qr[32];
                        tmem_ld32(t_lane + static_cast<uint32_t>(grp * 128 + half * 64 + hh * 32), qr);
                        tmem_ld_wait();
                        if (hh == 1) {      // Q accumulator fully read: K may issue the next Q tile of this group
                            tc_fence_before();
                            named_bar_arrive(kBarAfree + grp, kAfree);
                        }
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
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            sts_128u(qrow + static_cast<uint32_t>(((hh * 4 + u) ^ sw) << 4),
                                     pack_bf16x2(e[8 * u] * inv, e[8 * u + 1] * inv), pack_bf16x2(e[8 * u + 2] * inv, e[8 * u + 3] * inv),
                                     pack_bf16x2(e[8 * u + 4] * inv, e[8 * u + 5] * inv), pack_bf16x2(e[8 * u + 6] * inv, e[8 * u + 7] * inv));
                    }
                }
                TR(43, t);
                arrive(kBarEdone + grp, kPass2Edone);
                TR(44, t);
                // Y epilogue of the same tile: the Y MMA is issued as soon as the control warp sees the arrival above
                if (leader) mbar_wait(&bars->yfull[grp], ph_y);
                ph_y ^= 1u;
                publish();
                TR(45, t);
                {
                    constexpr int kCols = C / 2;                      // columns of this thread: half * kCols .. +kCols-1
                    const uint32_t yacc = t_lane + 256u + static_cast<uint32_t>(grp * 128 + half * kCols);
                    // sweep 1: sum of squares of (acc + bias) over this thread's columns (the accumulator is read again below:
                    // TMEM reads are cheap, 64 live fp32 values per thread at C = 128 are not)
                    float s0 = 0.0f, s1 = 0.0f;
                    uint32_t yr[32];          // C = 64: the thread's 32 columns stay here for sweep 2
#pragma unroll
                    for (int c32 = 0; c32 < kCols / 32; ++c32) {
                        tmem_ld32(yacc + static_cast<uint32_t>(c32 * 32), yr);
                        const float4* bp = reinterpret_cast<const float4*>(bias_s + half * kCols + c32 * 32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 bb = bp[i];
                            const float a0 = __uint_as_float(yr[4 * i]) + bb.x, a1 = __uint_as_float(yr[4 * i + 1]) + bb.y;
                            const float a2 = __uint_as_float(yr[4 * i + 2]) + bb.z, a3 = __uint_as_float(yr[4 * i + 3]) + bb.w;
                            s0 = fmaf(a0, a0, s0); s1 = fmaf(a1, a1, s1); s0 = fmaf(a2, a2, s0); s1 = fmaf(a3, a3, s1);
                        }
                    }
                    float* red = red_s + grp * 256;
                    red[half * 128 + row] = s0 + s1;
                    named_bar_sync(bar_g, kGroupThreads);
                    TR(46, t);
                    const float rinv = 1.0f / fmaxf(sqrtf(red[row] + red[128 + row]), 1e-12f);
                    // sweep 2: normalise, gain, + residual x (own row of the tile); written in place over the x tile.
                    // column c lives in atom c / 64, 16-byte unit (c % 64) / 8
                    const uint32_t xrow = sb + L::off_x + slot * L::kXTile + row * 128;
#pragma unroll
                    for (int c32 = 0; c32 < kCols / 32; ++c32) {
                        const int c0 = half * kCols + c32 * 32;
                        if (kCols > 32) tmem_ld32(yacc + static_cast<uint32_t>(c32 * 32), yr);
                        const uint32_t xa = xrow + static_cast<uint32_t>((c0 >> 6) * (kTileTok * 128));
                        uint4 xr[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) xr[u] = lds_128u(xa + static_cast<uint32_t>(((((c0 & 63) >> 3) + u) ^ sw) << 4));
                        const float4* bp = reinterpret_cast<const float4*>(bias_s + c0);
                        const float4* gp = reinterpret_cast<const float4*>(g_s + c0);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 b0 = bp[2 * u], b1 = bp[2 * u + 1], g0 = gp[2 * u], g1 = gp[2 * u + 1];
                            const uint32_t* yy = yr + 8 * u;
                            sts_128u(xa + static_cast<uint32_t>(((((c0 & 63) >> 3) + u) ^ sw) << 4),
                                     pack_bf16x2(fmaf((__uint_as_float(yy[0]) + b0.x) * rinv, g0.x, bf16_lo(xr[u].x)),
                                                 fmaf((__uint_as_float(yy[1]) + b0.y) * rinv, g0.y, bf16_hi(xr[u].x))),
                                     pack_bf16x2(fmaf((__uint_as_float(yy[2]) + b0.z) * rinv, g0.z, bf16_lo(xr[u].y)),
                                                 fmaf((__uint_as_float(yy[3]) + b0.w) * rinv, g0.w, bf16_hi(xr[u].y))),
                                     pack_bf16x2(fmaf((__uint_as_float(yy[4]) + b1.x) * rinv, g1.x, bf16_lo(xr[u].z)),
                                                 fmaf((__uint_as_float(yy[5]) + b1.y) * rinv, g1.y, bf16_hi(xr[u].z))),
                                     pack_bf16x2(fmaf((__uint_as_float(yy[6]) + b1.z) * rinv, g1.z, bf16_lo(xr[u].w)),
                                                 fmaf((__uint_as_float(yy[7]) + b1.w) * rinv, g1.w, bf16_hi(xr[u].w))));
                        }
                    }
                }
                arrive(kBarYdone + (t & 3), kGroupThreads + 32);
                TR(47, t);
                if (t + 2 < T) {        // own next tile: its norms now (the ring has had a whole tile time to deliver it)
                    poll_x(G2 + t + 2);
                    publish();
                    rn_compute(G2 + t + 2, 0, kTileTok);      // published by the barrier that opens that tile
                }
            }
            // both groups leave the pass together: the next image's first chunk epilogue reuses the union region / rn slots
            named_bar_sync(kBarAll, kEpiThreads);
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
    int r = static_cast<int>(cudaFuncSetAttribute(linattn_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, LaSmem<64>::kTotal + 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(linattn_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, LaSmem<128>::kTotal + 1024));
    return r;
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
    return (C == 64 || C == 128) && heads == 4 && d == 32 && n >= kTileTok && (n % kTileTok) == 0 && n_mem >= 0 && n_mem <= 4;
}

void launch_linattn_fused(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmWqkv, const CUtensorMap& tmWout,
                          const float* bias_out, const float* g_out, const float* mem_kv, const float* k_shift, int B, int n, int C,
                          int n_mem, int num_sms, int trace, bool pdl, cudaStream_t s) {
    LaParams p;
    p.bias_out = bias_out; p.g_out = g_out; p.mem_kv = mem_kv; p.k_shift = k_shift;
    p.B = B; p.n = n; p.n_mem = n_mem;
    p.trace = trace;
    const int grid = B < num_sms ? B : num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (C == 64 ? LaSmem<64>::kTotal : LaSmem<128>::kTotal) + 1024;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (C == 64) cudaLaunchKernelEx(&cfg, linattn_fused_kernel<64>, tmX, tmY, tmWqkv, tmWout, p);
    else cudaLaunchKernelEx(&cfg, linattn_fused_kernel<128>, tmX, tmY, tmWqkv, tmWout, p);
}

}  // namespace ddm
