// See conv_tc.cuh for the design.  sm_100a only: tcgen05.mma / TMEM / TMA.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace ddm {

namespace {

struct alignas(8) ConvBarriers {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t acc_full[4];
    uint64_t acc_empty[4];
    uint64_t res_full[4];        // lean epilogue: residual tile landed in group g's staging buffer
    uint64_t w_full;
    uint64_t xch_full[2];        // pair_n: the peer CTA's row sums of squares of tile parity i have landed in xch[i]
    uint32_t tmem_base;
    int issued;                  // MMA issue token: number of pipeline stages whose MMAs have all been issued
};

// Profiling / bisection switches of DDM_CONV_DEBUG that act INSIDE the kernel (skip epilogue / MMA / A loads, event trace,
// fence bisection) exist only in a build with -DDDM_CONV_DEBUG_BUILD: in the product build kDbg is 0, so every test below
// is a compile-time false and neither the branches nor the trace stores are in the SASS.  (Mode selection on the host --
// folding, epilogue groups, issuer modes -- keeps working either way.)
#ifdef DDM_CONV_DEBUG_BUILD
constexpr int kDbg = -1;
#else
constexpr int kDbg = 0;
#endif

// Device-side event trace for pipeline debugging (DDM_CONV_DEBUG & 128): CTA 0 records (role, event, tile, clock).
constexpr int kTraceRoles = 5, kTraceCap = 1024;      // per-role rings, no atomics: stores are fire-and-forget
__device__ long long g_trace[kTraceRoles * kTraceCap * 2];
__device__ __forceinline__ void trace_ev(bool on, int role, int ev, int idx, int& n) {
    if (!on || n >= kTraceCap) return;
    long long* dst = g_trace + (static_cast<size_t>(role) * kTraceCap + n) * 2;
    dst[0] = (static_cast<long long>(role) << 48) | (static_cast<long long>(ev) << 32) | static_cast<unsigned>(idx);
    dst[1] = clock64();
    ++n;
}

constexpr int kMaxParts = 4;                       // column parts per accumulator row (epilogue warps / 4)
constexpr int kBarPre = 1, kBarPost = 2, kBarRes = 3;   // named barriers of the generic epilogue's warps

// SiLU(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2 : one MUFU op per element
__device__ __forceinline__ float silu_f(float v) {
    const float h = 0.5f * v;
    return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

struct SmemPlan {
    int a_bytes, b_chunk_bytes, stage_bytes, wres_off, staging_off, colp_off, head_off, red_off, xch_off, bars_off, total;
};
__host__ __device__ inline SmemPlan make_plan(const ConvParams& p, int num_stages) {
    SmemPlan s;
    s.a_bytes = p.a_rows * (kChunkK * 2);
    s.b_chunk_bytes = p.block_n * (kChunkK * 2);
    s.stage_bytes = s.a_bytes + (p.b_resident ? 0 : p.n_dy * s.b_chunk_bytes);
    s.wres_off = num_stages * s.stage_bytes;
    s.staging_off = s.wres_off + (p.b_resident ? (p.k_chunks + p.res_chunks) * s.b_chunk_bytes : 0);
    s.colp_off = s.staging_off + (p.tma_store ? (p.staging_bufs > 1 ? p.staging_bufs : 1) * kTileM * p.block_n * 2 : 0);
    s.head_off = s.colp_off + (p.res_chunks ? 4 : 3) * p.n_pad * 4;     // fused head: weights [head_n][n_pad], then [2 groups][128] float4
    s.red_off = s.head_off + (p.head_n ? p.head_n * p.n_pad * 4 + 2 * kTileM * 16 : 0);
    s.xch_off = s.red_off + (p.rnorm_out != nullptr ? 2 : 1) * kMaxParts * kTileM * 4;     // red_b only with rnorm_out
    s.bars_off = s.xch_off + (p.pair_n ? 2 * kTileM * 4 : 0);
    s.total = s.bars_off + static_cast<int>(sizeof(ConvBarriers));
    return s;
}

// kEpiWarps epilogue warps (multiple of 4).  FAST: lean epilogue for the common case (bf16 output through smem
// staging + TMA store, full tiles where per-pixel side inputs are used, scale/shift shared by the batch); packed
// f32x2 arithmetic, all per-pixel address math hoisted out of the tile loop.
template <int kEpiWarps, bool FAST, int FOLD, int GROUPS, bool SPLITK = false, bool HEAD = false>
__global__ void __launch_bounds__(96 + 32 * kEpiWarps, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmR1,
               const __grid_constant__ ConvParams p) {
    constexpr int kEpiThreads = 32 * kEpiWarps;
    constexpr int kParts = kEpiWarps / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    if (p.tight_smem && smem != smem_raw) __trap();        // the host did not reserve the alignment slack for this plan
    const SmemPlan plan = make_plan(p, p.num_stages);
    const int stage_bytes = plan.stage_bytes;
    uint8_t* wres = smem + plan.wres_off;
    uint8_t* staging = smem + plan.staging_off;
    float* col_bias = reinterpret_cast<float*>(smem + plan.colp_off);
    float* col_mul = col_bias + p.n_pad;
    float* col_add = col_mul + p.n_pad;
    [[maybe_unused]] float* col_rbias = col_add + p.n_pad;              // fused shortcut only
    float* red_a = reinterpret_cast<float*>(smem + plan.red_off);      // [parts][128] partial sum of squares (pre-norm)
    float* red_b = red_a + kMaxParts * kTileM;                            // [parts][128] partial sum of squares (stored row)
    ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(smem + plan.bars_off);

    // shuffled from lane 0 so that the compiler knows the role dispatch is warp-uniform (uniform registers, no
    // vector->uniform election loops around every TMA / MMA instruction)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const bool ss_uniform = (p.scale_shift != nullptr) && (p.ss_stride == 0);
    const bool ss_batched = (p.scale_shift != nullptr) && (p.ss_stride != 0);
    const bool affine = (p.norm_g != nullptr) || ss_uniform;

    griddep_launch();                                       // the next kernel of the stream may start its own prologue
    const bool mcast = p.cluster == 2 && !p.pair_n;         // weights multicast inside CTA pairs (off by default, see api.cu)
    const uint32_t cl_count = mcast ? 2u : 1u;              // a stage is released by the issuers of every CTA of the cluster
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], cl_count);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(&bars->acc_full[a], p.issue_mode == 2 ? 1 : 2);   // per issuer thread, or (mode 2) the tile's owner only
            mbar_init(&bars->acc_empty[a], FAST ? kEpiWarps / GROUPS : kEpiWarps);   // FAST: one epilogue group per stage
        }
        mbar_init(&bars->w_full, 1);
        mbar_init(&bars->xch_full[0], kTileM);
        mbar_init(&bars->xch_full[1], kTileM);
        for (int g = 0; g < 4; ++g) mbar_init(&bars->res_full[g], 1);
        bars->issued = 0;
        fence_barrier_init();
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmOut);
        prefetch_tmap(&tmRes);
        if (p.res_chunks) prefetch_tmap(&tmR1);
        if (p.b_resident) {       // the weights of this (single) N tile are loaded once per CTA; they do not depend on the
                                  // previous kernel, so the load is in flight before griddep_wait()
            const SmemPlan pl = make_plan(p, p.num_stages);
            const int cpt = p.chunks0 + p.chunks1;
            mbar_arrive_expect_tx(&bars->w_full, static_cast<uint32_t>((p.k_chunks + p.res_chunks) * pl.b_chunk_bytes));
            for (int kc = 0; kc < p.k_chunks + p.res_chunks; ++kc) {       // (the shortcut's chunks follow the conv's along K)
                int slot = kc;
                if constexpr (FOLD) {       // stacked layout: block (dy, chunk) = [dx=-1 | dx=0 | dx=+1] x 64 rows
                    const int tap = kc / cpt, c = kc - tap * cpt;
                    slot = (p.fold_dyi[tap] * cpt + c) * 3 + p.fold_dxi[tap];
                }
                tma_load_2d(smem + pl.wres_off + slot * pl.b_chunk_bytes, &tmW, &bars->w_full, kc * kChunkK, 0);
            }
        }
    }
    if (warp == 1) {
        tmem_alloc(&bars->tmem_base, static_cast<uint32_t>(p.tmem_cols));
        tmem_relinquish();
    }
    griddep_wait();               // everything below reads what earlier kernels of the stream wrote (activations, scale/shift row)
    // per-column epilogue vectors: v = acc*rs + bias ; [v *= rinv] ; v = v*mul + add
    for (int n = threadIdx.x; n < p.n_pad; n += blockDim.x) {
        const bool in = n < p.N;
        float m = 1.0f, a = 0.0f;
        if (in && p.norm_g != nullptr) m = __ldg(p.norm_g + n);
        if (in && ss_uniform) {
            m *= __ldg(p.scale_shift + n) + 1.0f;
            a = __ldg(p.scale_shift + p.N + n);
        }
        if (FAST && p.act == 1 && affine) { m *= 0.5f; a *= 0.5f; }   // SiLU(z) = h + h tanh(h), h = z/2: fold the 1/2
        col_bias[n] = (in && p.bias != nullptr) ? __ldg(p.bias + n) : 0.0f;
        col_mul[n] = m;
        col_add[n] = a;
        if (p.res_chunks) col_rbias[n] = (in && p.rbias != nullptr) ? __ldg(p.rbias + n) : 0.0f;
    }
    if constexpr (HEAD) {
        float* hw = reinterpret_cast<float*>(smem + plan.head_off);
        for (int i = threadIdx.x; i < p.head_n * p.n_pad; i += blockDim.x) {
            const int o = i / p.n_pad, n = i - o * p.n_pad;
            hw[i] = n < p.N ? __ldg(p.head_w + o * p.N + n) : 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster == 2) cluster_sync_all();      // the peer's barriers must be initialised before anything is multicast
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const int chunks_per_tap = p.chunks0 + p.chunks1;
    const bool tr = ((p.debug & kDbg) & 128) && blockIdx.x == 0;
    const bool tr_iss = ((p.debug & kDbg) & (128 | 4096)) && blockIdx.x == 0;      // 4096: only the issuers' token/issued events
    const bool tr_tile = ((p.debug & kDbg) & (128 | 256)) && blockIdx.x == 0;     // 256: per-tile events only (unperturbed timing)
    int trn = 0;
    // Tile sequence of this CTA.  Without clusters: tiles blockIdx.x, +gridDim.x, ...  With 2-CTA clusters (weights
    // multicast): cluster c takes pair sequence c, c + n_clusters, ...; the pair's two M tiles go to the two CTAs and
    // share the N tile.  An odd tile count leaves a phantom tile past the batch end: all its loads are zero-filled
    // and all its stores clipped, but it keeps the cluster's barrier protocol in lockstep.
    const int cl = p.cluster;
    const int cl_rank = cl == 2 ? static_cast<int>(blockIdx.x & 1) : 0;
    auto seq_tile = [&](int q, int& n_tile, int& m_tile) -> bool {
        if (cl == 2 && p.pair_n) {      // cluster c: M tiles c, c + n_clusters, ..; the CTA's rank is its N tile
            const int ct = static_cast<int>(blockIdx.x >> 1) + q * static_cast<int>(gridDim.x >> 1);
            if (ct >= p.m_tiles) return false;
            m_tile = ct;
            n_tile = cl_rank;
        } else if (cl == 2) {
            const int ct = static_cast<int>(blockIdx.x >> 1) + q * static_cast<int>(gridDim.x >> 1);
            if (ct >= p.pairs * p.n_tiles) return false;
            n_tile = ct >= p.pairs ? 1 : 0;
            m_tile = (ct - n_tile * p.pairs) * 2 + cl_rank;
        } else {
            int t = static_cast<int>(blockIdx.x) + q * static_cast<int>(gridDim.x);
            if (t >= p.total_tiles) return false;
            if (SPLITK && p.ksplit > 1) {       // t = ((m_tile * ksplit) + ks) * n_tiles + n_tile
                n_tile = t % p.n_tiles;
                m_tile = (t / p.n_tiles) / p.ksplit;
                return true;
            }
            // N tile fastest: the (at most two) N tiles of an M tile run at the same time on neighbouring CTAs, so the
            // activations are fetched from HBM once and from L2 afterwards (C_out = 384: 12 % of the traffic)
            if (p.n_tiles == 1) { m_tile = t; n_tile = 0; }
            else if (p.n_tiles == 2) { m_tile = t >> 1; n_tile = t & 1; }
            else { m_tile = t / p.n_tiles; n_tile = t - m_tile * p.n_tiles; }
        }
        return true;
    };

    // split-K: pipeline-stage range [lo, hi) of the q-th tile of this CTA (all stages without a split)
    auto stage_range = [&](int q, int& lo, int& hi, int& ks) {
        lo = 0; hi = 0x7FFFFFFF; ks = 0;
        if (p.ksplit > 1) {
            const int t = static_cast<int>(blockIdx.x) + q * static_cast<int>(gridDim.x);
            ks = (t / p.n_tiles) % p.ksplit;
            lo = ks * p.ks_per;
            hi = lo + p.ks_per;
        }
    };
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        {
            // (the resident weights were requested by thread 0 in the prologue, ahead of griddep_wait)
            // issue_mode 2 splits the ring in two halves, one per issuer thread (tile parity): every mbarrier then has
            // a single consumer that visits it in lap order, which parity-only waits require.
            const bool split = p.issue_mode == 2;
            const int ring_stages = split ? (p.num_stages >> 1) : p.num_stages;
            int stage_r[2] = {0, 0};
            uint32_t phase_r[2] = {0u, 0u};
            int n_tile, m_tile;
            for (int q = 0; seq_tile(q, n_tile, m_tile); ++q) {
                const int ring = split ? (q & 1) : 0;
                const int base = ring * ring_stages;
                int stage = stage_r[ring];
                uint32_t phase = phase_r[ring];
                const int tx = m_tile % p.tiles_x;
                const int ty = (m_tile / p.tiles_x) % p.tiles_y;
                const int tb = m_tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
                const int n0 = n_tile * p.block_n;
                [[maybe_unused]] int st_lo = 0, st_hi = 0, ks_unused = 0, si = 0;
                if constexpr (SPLITK) stage_range(q, st_lo, st_hi, ks_unused);       // (split-K: its own instantiation of the generic kernel)
                for (int s = 0; s < p.n_slabs; ++s) {
                    const int cx = x0 + p.slab_dx[s], cy = y0 + p.slab_dy0[s], cp = p.slab_p[s];
                    for (int c = 0; c < chunks_per_tap; ++c) {
                        if constexpr (SPLITK) {
                            const int sidx = si++;
                            if (sidx < st_lo || sidx >= st_hi) continue;        // another split's stage
                        }
                        mbar_wait(&bars->empty[base + stage], phase ^ 1u);
                        if (elect_one()) {
                            trace_ev(tr, 0, 0, q, trn);
                            uint8_t* a_dst = smem + (base + stage) * stage_bytes;
                            const bool skip_a = ((p.debug & kDbg) & 4) != 0;      // profiling: no A traffic
                            const uint32_t tx_bytes = static_cast<uint32_t>(skip_a ? stage_bytes - plan.a_bytes : stage_bytes);
                            mbar_arrive_expect_tx(&bars->full[base + stage], tx_bytes);
                            if (skip_a) {
                            } else if (c < p.chunks0) {
                                tma_load_5d(a_dst, &tmA0, &bars->full[base + stage], c * kChunkK, cx, cp, cy, b0);
                            } else {
                                tma_load_5d(a_dst, &tmA1, &bars->full[base + stage], (c - p.chunks0) * kChunkK, cx, cp, cy, b0);
                            }
                            if (!p.b_resident) {
                                for (int j = 0; j < p.n_dy; ++j) {
                                    uint8_t* b_dst = a_dst + plan.a_bytes + j * plan.b_chunk_bytes;
                                    const int kcol = (p.slab_tap[s][j] * chunks_per_tap + c) * kChunkK;
                                    if (mcast) {        // each CTA fetches half of the rows and multicasts them to both
                                        const int half_rows = p.block_n >> 1;
                                        tma_load_2d_mc(b_dst + cl_rank * half_rows * (kChunkK * 2), &tmW, &bars->full[base + stage], kcol,
                                                       n0 + cl_rank * half_rows, 0x3);
                                    } else {
                                        tma_load_2d(b_dst, &tmW, &bars->full[base + stage], kcol, n0);
                                    }
                                }
                            }
                            trace_ev(tr, 0, 1, q, trn);
                        }
                        __syncwarp();
                        if (++stage == ring_stages) { stage = 0; phase ^= 1u; }
                    }
                }
                for (int c = 0; c < p.res_chunks; ++c) {       // fused shortcut: the (unshifted) 128-pixel tiles of its sources
                    mbar_wait(&bars->empty[base + stage], phase ^ 1u);
                    if (elect_one()) {
                        uint8_t* a_dst = smem + (base + stage) * stage_bytes;
                        mbar_arrive_expect_tx(&bars->full[base + stage], static_cast<uint32_t>(kATileBytes + (p.b_resident ? 0 : plan.b_chunk_bytes)));
                        if (c < p.res_chunks0) tma_load_5d(a_dst, &tmRes, &bars->full[base + stage], c * kChunkK, x0, 0, y0, b0);
                        else tma_load_5d(a_dst, &tmR1, &bars->full[base + stage], (c - p.res_chunks0) * kChunkK, x0, 0, y0, b0);
                        if (!p.b_resident)       // streamed weights: the shortcut's chunk rides in the stage's first weight slot
                            tma_load_2d(a_dst + plan.a_bytes, &tmW, &bars->full[base + stage], (p.k_chunks + c) * kChunkK, n0);
                    }
                    __syncwarp();
                    if (++stage == ring_stages) { stage = 0; phase ^= 1u; }
                }
                stage_r[ring] = stage;
                phase_r[ring] = phase;
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------ MMA issuers (two threads, ordered alternation)
        // One tcgen05.mma costs its issuing thread ~50 cycles, the pipe accepts only a few MMAs ahead, and a stage
        // hand-over (mbarrier try_wait ~175 cycles even when complete, + tcgen05.commit ~60) is pure dead time for a
        // single issuer (scripts/micro/*.cu).  Two issuer threads therefore take alternate pipeline stages: while one
        // is blocked feeding the pipe, the other performs its hand-over.  To keep results bit-reproducible the MMAs
        // must still enter the pipe in stage order: an issuer only starts stage g once `issued` says that stage g-1
        // has been completely issued (a shared-memory token, ~30 cycle poll instead of an mbarrier round trip).
        {
            const int me = warp - 1;
            const bool dual = p.issue_mode == 1;
            // FOLD: p.fold = 3 -> one slab, N = 3*64 (all dx groups); p.fold = 2 -> slab 0 (dx = 0) feeds the [dx=-1 | dx=0]
            // groups with N = 128 and slab 1 (dx = +1) accumulates straight into the dx = 0 group with N = 64
            const uint32_t idesc = umma_idesc_bf16(kTileM, static_cast<uint32_t>(FOLD ? FOLD * p.block_n : p.block_n));
            const uint32_t idesc_n64 = umma_idesc_bf16(kTileM, static_cast<uint32_t>(p.block_n));
            // Only the start-address field (bits 0-13, address >> 4) of the smem descriptors changes: build the
            // constant part once and add precomputed 16-byte-unit offsets.
            const uint64_t desc_base = umma_desc_sw128(0);
            const uint32_t smem_lo = smem_u32(smem) >> 4;
            const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
            const uint32_t dy_step = static_cast<uint32_t>(p.bw * (kChunkK * 2)) >> 4;
            const uint32_t bchunk_step = static_cast<uint32_t>(plan.b_chunk_bytes) >> 4;
            const uint32_t b_in_stage = static_cast<uint32_t>(plan.a_bytes) >> 4;
            const uint32_t wres_lo = smem_u32(wres) >> 4;
            const int n_dy = p.n_dy;
            const bool resident = p.b_resident != 0;
            const bool do_mma = ((p.debug & kDbg) & 2) == 0;
            volatile int* issued = &bars->issued;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int g = 0;                      // global stage counter (same sequence in both issuers)
            if (resident) mbar_wait(&bars->w_full, 0);
            int n_tile_unused, m_tile_unused;
            const bool by_tile = p.issue_mode == 2;
            const int ring_stages = by_tile ? (p.num_stages >> 1) : p.num_stages;
            const int ring_base = by_tile ? me * ring_stages : 0;
            const int stages_per_tile = p.n_slabs * chunks_per_tap + p.res_chunks;
            for (int q = 0; seq_tile(q, n_tile_unused, m_tile_unused); ++q) {
                if (by_tile && (q & 1) != me) {
                    // issue_mode 2: the issuers take alternate TILES.  Consecutive tiles use different accumulators,
                    // so no ordering between the two threads is needed for reproducible sums (each accumulator is
                    // fed by one thread in program order), and while one thread sits in a barrier hand-over the
                    // other one's MMAs keep the pipe busy.  Each thread owns one half of the smem ring: sharing one
                    // ring would make a thread wait on barriers it last saw several phases ago, and mbarrier waits
                    // only know the phase parity (both an arithmetic skip and a walk with non-consuming waits were
                    // tried; they alias / dead-lock when one thread runs a lap ahead of the other).
                    g += stages_per_tile;
                    if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
                    continue;
                }
                mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1u);
                if (tr_tile && lane == 0) trace_ev(tr_tile, 1 + me, 0, q, trn);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
                bool first = true;              // first stage of the tile: its first MMA overwrites the accumulator
                bool mine_any = false;
                [[maybe_unused]] int st_lo = 0, st_hi = 0, ks_unused = 0, si = 0;
                if constexpr (SPLITK) stage_range(q, st_lo, st_hi, ks_unused);
                for (int s = 0; s < p.n_slabs; ++s) {
                    const uint32_t t0 = static_cast<uint32_t>(p.slab_tap[s][0] * chunks_per_tap) * bchunk_step;
                    const uint32_t t1 = static_cast<uint32_t>(p.slab_tap[s][1] * chunks_per_tap) * bchunk_step;
                    const uint32_t t2 = static_cast<uint32_t>(p.slab_tap[s][2] * chunks_per_tap) * bchunk_step;
                    for (int c = 0; c < chunks_per_tap; ++c) {
                        if constexpr (SPLITK) {
                            const int sidx = si++;
                            if (sidx < st_lo || sidx >= st_hi) continue;        // another split's stage
                        }
                        const bool mine = by_tile ? true : (dual ? ((g & 1) == me) : (me == 0));
                        if (mine) {
                            mbar_wait(&bars->full[ring_base + stage], phase);
                            if (tr && lane == 0) trace_ev(tr, 1 + me, 1, g, trn);
                            tc_fence_after();
                            if (tr_iss && lane == 0) trace_ev(tr_iss, 1 + me, 2, g, trn);
                            const uint32_t a_lo = smem_lo + static_cast<uint32_t>(ring_base + stage) * stage_step;
                            const uint32_t bres = wres_lo + static_cast<uint32_t>(c) * bchunk_step;
                            if (elect_one()) {
                            if (dual) {         // wait for the token: every MMA of stage g-1 has been issued.  Polled by the
                                uint32_t spins = 0;     // issuing lane itself, with everything else already set up
                                while (*issued < g) { if (++spins > (1u << 28)) __trap(); }
                            }
                            if (do_mma) {
                                uint32_t accumulate = first ? 0u : 1u;
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    if (j < n_dy) {
                                        // dy shift = j image rows = j * bw * 128 bytes (1024B-aligned) into the slab
                                        const uint64_t a_desc = desc_base | static_cast<uint64_t>(a_lo + j * dy_step);
                                        const uint32_t b_lo =
                                            FOLD ? wres_lo + static_cast<uint32_t>((j * chunks_per_tap + c) * 3 + (s ? 2 : 0)) * bchunk_step
                                                 : (resident ? bres + (j == 0 ? t0 : (j == 1 ? t1 : t2))
                                                             : a_lo + b_in_stage + j * bchunk_step);
                                        const uint64_t b_desc = desc_base | static_cast<uint64_t>(b_lo);
#pragma unroll
                                        for (int k = 0; k < kChunkK / 16; ++k) {
                                            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 units
                                            if (FOLD && s) umma_bf16(d_tmem + p.block_n, a_desc + 2u * k, b_desc + 2u * k, idesc_n64, 1u);
                                            else umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, accumulate);
                                            accumulate = 1u;
                                        }
                                    }
                                }
                            }
                            trace_ev(tr_iss, 1 + me, 3, g, trn);
                            if (dual) { if (!((p.debug & kDbg) & 8192)) __threadfence_block(); *issued = g + 1; }     // pass the token
                            if (mcast) umma_commit_mc(&bars->empty[ring_base + stage], 0x3); else umma_commit(&bars->empty[ring_base + stage]);
                            trace_ev(tr, 1 + me, 4, g, trn);
                            }
                            __syncwarp();
                            mine_any = true;
                        }
                        first = false;
                        ++g;
                        if (++stage == ring_stages) { stage = 0; phase ^= 1u; }
                    }
                }
                for (int c = 0; c < p.res_chunks; ++c) {       // fused shortcut: same hand-over protocol, second accumulator
                    const bool mine = by_tile ? true : (dual ? ((g & 1) == me) : (me == 0));
                    if (mine) {
                        mbar_wait(&bars->full[ring_base + stage], phase);
                        tc_fence_after();
                        const uint32_t a_lo = smem_lo + static_cast<uint32_t>(ring_base + stage) * stage_step;
                        const uint32_t b_lo = resident ? wres_lo + static_cast<uint32_t>(p.k_chunks + c) * bchunk_step : a_lo + b_in_stage;
                        if (elect_one()) {
                            if (dual) {
                                uint32_t spins = 0;
                                while (*issued < g) { if (++spins > (1u << 28)) __trap(); }
                            }
                            if (do_mma) {
                                const uint64_t a_desc = desc_base | static_cast<uint64_t>(a_lo);
                                const uint64_t b_desc = desc_base | static_cast<uint64_t>(b_lo);
#pragma unroll
                                for (int k = 0; k < kChunkK / 16; ++k)
                                    umma_bf16(d_tmem + p.block_n, a_desc + 2u * k, b_desc + 2u * k, idesc_n64, (c | k) ? 1u : 0u);
                            }
                            if (dual) { __threadfence_block(); *issued = g + 1; }
                            umma_commit(&bars->empty[ring_base + stage]);
                        }
                        __syncwarp();
                        mine_any = true;
                    }
                    ++g;
                    if (++stage == ring_stages) { stage = 0; phase ^= 1u; }
                }
                // acc_full expects one arrival per issuer: after this thread's MMAs retire, or at once if it had none
                if (elect_one()) {
                    if (mine_any) umma_commit(&bars->acc_full[acc]); else if (!by_tile) mbar_arrive(&bars->acc_full[acc]);
                }
                __syncwarp();
                if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if constexpr (FAST) {
        // ------------------------------------------------------------------ lean epilogue (warps 3..3+kEpiWarps-1)
        // The critical path of a tile's epilogue is ONE warp's serial instruction stream (~5 cycles per dependent
        // instruction with few warps per scheduler), so: many warps with little work each (16 warps -> one 16-column
        // chunk per thread for C_out = 64), no per-pixel address math unless a side input needs it, the first chunk
        // stays in registers between the norm pass and the store pass, packed f32x2 arithmetic.
        // Two independent epilogue groups of kEpiWarps/2 warps: group g owns accumulator stage g, staging buffer g and
        // its own named barriers, and handles every other tile of this CTA, so two tile epilogues are in flight.
        // GROUPS = 2: 8 warps per tile (thread = pixel x column half, sums of squares meet in shared memory).
        // GROUPS = 4: 4 warps per tile (thread = pixel x all columns: no cross-warp reduction, four tiles in flight on
        // the four accumulator stages, half as many threads per named barrier).
        constexpr int kGroupWarps = kEpiWarps / GROUPS;
        constexpr int kGroupThreads = 32 * kGroupWarps;
        constexpr int kGParts = kGroupWarps / 4;
        const int grp = (warp - 3) / kGroupWarps;
        const int ew = (warp - 3) - grp * kGroupWarps;
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int part = ew >> 2;               // column part (within the group) handled by this warp
        const int bar0 = 1 + grp * 3;           // this group's named barriers: bar0 + {0 pre, 1 post, 2 residual / accumulator}
        const int r = q * 32 + lane;            // accumulator row == tile pixel
        const bool leader_warp = (ew == 0);
        const bool store_leader = leader_warp && (lane == 0);
        const bool has_norm = p.norm_g != nullptr;
        const bool has_act = p.act == 1;
        const bool act_prescaled = has_act && affine;
        const bool res_smem = p.residual != nullptr;
        const bool res_tma = res_smem && p.res_tma != 0;
        const bool fused_res = p.res_chunks != 0;      // the shortcut's accumulator (columns block_n ..) is added after the activation
        const uint32_t sb_rbias = smem_u32(col_rbias);
        uint32_t res_phase = 0;
        const bool want_rs = p.row_scale != nullptr, want_rn = p.rnorm_out != nullptr;
        const bool need_geo = leader_warp || (res_smem && !res_tma) || want_rs || want_rn || HEAD;   // who needs the tile's coordinates
        const bool skip = ((p.debug & kDbg) & 1) != 0;   // profiling: no epilogue math / stores
        const int stg_bytes = kTileM * p.block_n * 2;
        const int tiles_xy = p.tiles_x * p.tiles_y;
        // offset of this thread's pixel inside a (full) tile of the [B,H,W] grid, for row_scale / rnorm_out
        const int row_off = ((r >> (p.bw_shift + p.bh_shift)) * p.H + ((r >> p.bw_shift) & (p.bh - 1))) * p.W + (r & (p.bw - 1));
        const int sw = r & 7;
        constexpr bool fold3 = FOLD == 3;
        const bool has_left = (lane & (p.bw - 1)) != 0, has_right = (lane & (p.bw - 1)) != (p.bw - 1);   // FOLD (bw <= 32)
        const int et = threadIdx.x - 96 - grp * kGroupThreads;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

        struct TileGeo { int x0, y0, b0; };
        auto decode = [&](int m_tile) {
            TileGeo t;
            int tx, ty, tb;
            if (p.tiles_pow2) {
                tx = m_tile & (p.tiles_x - 1);
                ty = (m_tile >> p.tx_shift) & (p.tiles_y - 1);
                tb = m_tile >> (p.tx_shift + p.ty_shift);
            } else {
                tb = m_tile / tiles_xy;
                const int rem = m_tile - tb * tiles_xy;
                ty = rem / p.tiles_x;
                tx = rem - ty * p.tiles_x;
            }
            t.x0 = tx * p.bw; t.y0 = ty * p.bh; t.b0 = tb * p.bb;
            return t;
        };
        // residual tile -> staging buffer in the staging layout (coalesced 16-byte cp.async by all epilogue threads)
        auto fetch_residual = [&](int n_tile, int m_tile, uint8_t* buf) {
            const TileGeo t = decode(m_tile);
            const int n0 = n_tile * p.block_n;
            const int upr = min(p.block_n, p.N - n0) >> 3;       // 16-byte units per row
            for (int u = et; u < kTileM * upr; u += kGroupThreads) {
                const int row = u / upr, cu = u - row * upr;
                const int rx = t.x0 + (row & (p.bw - 1));
                const int ry = t.y0 + ((row >> p.bw_shift) & (p.bh - 1));
                const int rb = t.b0 + (row >> (p.bw_shift + p.bh_shift));
                uint8_t* dst = buf + (cu >> 3) * (kTileM * 128) + row * 128 + (((cu & 7) ^ (row & 7)) << 4);
                if (rx < p.W && ry < p.H && rb < p.B) {
                    const long long pix = (static_cast<long long>(rb) * p.OH + (ry * p.sy + p.oy)) * p.OW + (rx * p.sx + p.ox);
                    cp_async_16(dst, p.residual + pix * p.ld_res + n0 + cu * 8);
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
            }
            cp_async_commit();
        };
        // shared-space addresses of the per-column vectors, the reduction scratch and this thread's staging row
        const uint32_t sb_bias = smem_u32(col_bias), sb_mul = smem_u32(col_mul), sb_add = smem_u32(col_add);
        // chunk range of this thread inside an N tile (the last tile may be narrower)
        auto chunk_range = [&](int n_tile, int& lo, int& hi) {
            const int nc = min(p.block_n, p.N - n_tile * p.block_n);
            const int nch = nc > 0 ? (nc + 15) >> 4 : 0;
            const int per = (nch + kGParts - 1) / kGParts;
            lo = min(nch, part * per);
            hi = min(nch, lo + per);
        };
        const uint64_t half2 = pk2(0.5f, 0.5f);
        // tile sequence number q -> accumulator stage q % acc_stages (2 or 4; group g sees stages g, g + 2), phase (q / stages) & 1
        const int acc_mask = p.acc_stages - 1, acc_shift = p.acc_stages == 4 ? 2 : 1;
        uint8_t* const buf = staging + grp * stg_bytes;
        float* const gred_a = red_a + grp * kGParts * kTileM;
        float* const gred_b = red_b + grp * kGParts * kTileM;
        // measured: pays for the dx-folded kernels (-6 %), neutral or worse elsewhere (with two accumulator stages it
        // serialises the next tile's MMA with this tile's store)
        const bool merge_acc = FOLD != 0 && p.acc_stages == 4 && ((p.debug & kDbg) & 1048576) == 0;
        bool acc_ready = false;
        int n_tile, m_tile;
        for (int q = grp; seq_tile(q, n_tile, m_tile); q += GROUPS) {
            const bool real_tile = m_tile < p.m_tiles;          // false only for a cluster's phantom tile
            const int acc = q & acc_mask;
            const uint32_t acc_phase = static_cast<uint32_t>(q >> acc_shift) & 1u;
            const int n0 = n_tile * p.block_n;
            int c_lo, c_hi;
            chunk_range(n_tile, c_lo, c_hi);
            TileGeo tg = {0, 0, 0};
            int tile_pix = 0;
            uint64_t rs2 = pk2(1.0f, 1.0f);
            if (need_geo) {
                tg = decode(m_tile);
                if ((want_rs || want_rn) && real_tile) {
                    tile_pix = (tg.b0 * p.H + tg.y0) * p.W + tg.x0 + row_off;     // full tiles only (host-checked)
                    if (want_rs) { const float rs = __ldg(p.row_scale + tile_pix); rs2 = pk2(rs, rs); }
                }
            }
            if (res_tma) {      // one TMA load per 64-channel group, issued once the previous store has drained the buffer
                if (store_leader) {
                    bulk_wait_group_read<0>();
                    const int groups = (min(p.block_n, p.N - n0) + 63) >> 6;
                    mbar_arrive_expect_tx(&bars->res_full[grp], static_cast<uint32_t>(groups * kTileM * 128));
                    for (int g = 0; g < groups; ++g) {
                        const int ch = n0 + g * 64;
                        if (p.sy == 2) tma_load_5d(buf + g * (kTileM * 128), &tmRes, &bars->res_full[grp], p.ox * p.ld_res + ch, tg.x0, p.oy, tg.y0, tg.b0);
                        else tma_load_5d(buf + g * (kTileM * 128), &tmRes, &bars->res_full[grp], ch, tg.x0, 0, tg.y0, tg.b0);
                    }
                }
            } else if (res_smem) {     // the group's previous TMA store must have drained its staging buffer before the fetch
                if (store_leader) bulk_wait_group_read<0>();
                named_bar_sync(bar0 + 2, kGroupThreads);
                fetch_residual(n_tile, m_tile, buf);
            }

            // The leader polls the accumulator's mbarrier and the group learns of it through a named barrier (16 polling
            // warps slow down every other mbarrier operation of the CTA).  From the second tile on, that barrier is the
            // previous tile's store barrier: the leader waits for the NEXT accumulator just before it (see below).
            if (!acc_ready) {
                if (store_leader) { mbar_wait(&bars->acc_full[acc], acc_phase); trace_ev(tr_tile, 3 + grp, 0, q, trn); }
                named_bar_sync(bar0 + 2, kGroupThreads);
            }
            if (store_leader) trace_ev(tr, 3 + grp, 1, q, trn);
            tc_fence_after();
            const uint32_t t_row = t_lane + static_cast<uint32_t>(acc * p.acc_stride);
            const bool one_chunk = (c_hi - c_lo) == 1;
            // 16 accumulator columns of this thread's pixel.  FOLD: the three dx column groups, the outer two taken
            // from the x-1 / x+1 neighbour's lane (tile rows are whole image rows inside one warp; zero at the border)
            auto load_chunk = [&](int c, uint64_t (&v)[8]) {
                __syncwarp();
                if constexpr (FOLD) {
                    uint64_t t[8];
                    [[maybe_unused]] uint64_t u[8];
                    tmem_ld16x2(t_row + c * 16, t);                       // dx = -1 group
                    tmem_ld16x2(t_row + p.block_n + c * 16, v);           // dx = 0 group
                    tmem_ld_wait();
                    if constexpr (fold3) tmem_ld16x2(t_row + 2 * p.block_n + c * 16, u);   // dx = +1 group, in flight during the shuffles
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t a = __shfl_up_sync(0xffffffffu, static_cast<unsigned long long>(t[j]), 1);
                        if (has_left) v[j] = fadd2(v[j], a);
                    }
                    if constexpr (fold3) {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint64_t b = __shfl_down_sync(0xffffffffu, static_cast<unsigned long long>(u[j]), 1);
                            if (has_right) v[j] = fadd2(v[j], b);
                        }
                    }
                } else {
                    tmem_ld16x2(t_row + c * 16, v);
                    tmem_ld_wait();
                }
            };

            // ---- pass 1: bias (+ row scale), sum of squares; the first chunk (v0) stays in registers
            uint64_t v0[8];
            [[maybe_unused]] uint64_t v1[8];     // FOLD: the second chunk stays in registers as well
            bool have_v1 = false;
            const bool two_chunks = (c_hi - c_lo) == 2;
            bool released = false;
            if (c_lo < c_hi && !skip) {
                uint64_t s01 = 0ull, s23 = 0ull;
                load_chunk(c_lo, v0);
                {
                    const uint32_t ba = sb_bias + static_cast<uint32_t>(n0 + c_lo * 16) * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 bb = lds_128(ba + j * 16);
                        v0[2 * j] = ffma2(v0[2 * j], rs2, bb.x);
                        v0[2 * j + 1] = ffma2(v0[2 * j + 1], rs2, bb.y);
                        s01 = ffma2(v0[2 * j], v0[2 * j], s01);
                        s23 = ffma2(v0[2 * j + 1], v0[2 * j + 1], s23);    // padded columns have acc == 0 and bias == 0
                    }
                }
                if (has_norm) {
                    for (int c = c_lo + 1; c < c_hi; ++c) {
                        uint64_t v[8];
                        load_chunk(c, v);
                        const uint32_t ba = sb_bias + static_cast<uint32_t>(n0 + c * 16) * 4u;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const ulonglong2 bb = lds_128(ba + j * 16);
                            const uint64_t f0 = ffma2(v[2 * j], rs2, bb.x);
                            const uint64_t f1 = ffma2(v[2 * j + 1], rs2, bb.y);
                            s01 = ffma2(f0, f0, s01);
                            s23 = ffma2(f1, f1, s23);
                            if constexpr (FOLD != 0) { v1[2 * j] = f0; v1[2 * j + 1] = f1; }
                        }
                    }
                    if constexpr (FOLD != 0) have_v1 = two_chunks;     // re-assembling a folded chunk costs TMEM loads + shuffles
                    float a0, a1, a2, a3;
                    upk2(s01, a0, a1);
                    upk2(s23, a2, a3);
                    gred_a[part * kTileM + r] = (a0 + a1) + (a2 + a3);
                }
                if ((one_chunk || have_v1) && !fused_res) {   // nothing more to read from TMEM: release the accumulator right away
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
                    released = true;
                }
            }
            // the staging buffer about to be written must have been drained by its previous TMA store
            if (store_leader && !res_smem) bulk_wait_group_read<0>();
            if (res_tma) { if (store_leader) mbar_wait(&bars->res_full[grp], res_phase); res_phase ^= 1u; }
            else if (res_smem) cp_async_wait_all();
            if (store_leader) trace_ev(tr, 3 + grp, 2, q, trn);
            named_bar_sync(bar0, kGroupThreads);
            if (store_leader) trace_ev(tr, 3 + grp, 3, q, trn);
            uint64_t rinv2 = pk2(1.0f, 1.0f);
            if (has_norm && !skip) {
                float t = gred_a[r];
#pragma unroll
                for (int i = 1; i < kGParts; ++i) t += gred_a[i * kTileM + r];
                const float rinv = 1.0f / fmaxf(sqrtf(t), 1e-12f);
                rinv2 = pk2(rinv, rinv);
            }

            // ---- pass 2: normalise, scale/shift, SiLU, residual, bf16 -> staging
            [[maybe_unused]] uint64_t hacc[4] = {0ull, 0ull, 0ull, 0ull};        // HEAD: packed partial dot products of this thread's columns
            float out_sumsq = 0.0f;
            const uint32_t my_row = smem_u32(buf) + static_cast<uint32_t>(r) * 128u;
            for (int c = c_lo; c < c_hi && !skip; ++c) {
                uint64_t (&v)[8] = v0;          // the first chunk is already there; later chunks overwrite it (no copies)
                const int nb = n0 + c * 16;
                [[maybe_unused]] uint64_t rv[8];
                if (FOLD == 0 && fused_res) {   // in flight during the activation math below
                    __syncwarp();
                    tmem_ld16x2(t_row + p.block_n + c * 16, rv);
                }
                if (c == c_lo) {
                    if (FOLD == 0 && fused_res) tmem_ld_wait();
                } else if (FOLD != 0 && have_v1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = v1[j];
                } else {
                    load_chunk(c, v);
                    const uint32_t ba = sb_bias + static_cast<uint32_t>(nb) * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 bb = lds_128(ba + j * 16);
                        v[2 * j] = ffma2(v[2 * j], rs2, bb.x);
                        v[2 * j + 1] = ffma2(v[2 * j + 1], rs2, bb.y);
                    }
                }
                if (c == c_hi - 1 && !released) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);     // one arrival per epilogue warp
                    released = true;
                }
                if (affine) {
                    const uint32_t ma = sb_mul + static_cast<uint32_t>(nb) * 4u, aa_ = sb_add + static_cast<uint32_t>(nb) * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 mm = lds_128(ma + j * 16), aa = lds_128(aa_ + j * 16);
                        v[2 * j] = ffma2(fmul2(v[2 * j], rinv2), mm.x, aa.x);
                        v[2 * j + 1] = ffma2(fmul2(v[2 * j + 1], rinv2), mm.y, aa.y);
                    }
                }
                if (has_act) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t h = act_prescaled ? v[j] : fmul2(v[j], half2);
                        float h0, h1;
                        upk2(h, h0, h1);
                        v[j] = ffma2(h, pk2(tanh_approx(h0), tanh_approx(h1)), h);
                    }
                }
                const int cl = c * 16;                       // column inside this N tile
                const uint32_t rowp = my_row + static_cast<uint32_t>((cl >> 6) * (kTileM * 128));
                const int u = (cl & 63) >> 3;                // 16-byte unit inside the 128-byte row
                const uint32_t s0 = rowp + static_cast<uint32_t>(((u) ^ sw) << 4);
                const uint32_t s1 = rowp + static_cast<uint32_t>(((u + 1) ^ sw) << 4);
                if (FOLD == 0 && fused_res) {
                    const uint32_t ra = sb_rbias + static_cast<uint32_t>(nb) * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 bb = lds_128(ra + j * 16);
                        v[2 * j] = fadd2(v[2 * j], fadd2(rv[2 * j], bb.x));
                        v[2 * j + 1] = fadd2(v[2 * j + 1], fadd2(rv[2 * j + 1], bb.y));
                    }
                }
                if (res_smem) {
                    const uint4 r0 = lds_128u(s0), r1 = lds_128u(s1);
                    v[0] = fadd2(v[0], bf2_to_f2(r0.x)); v[1] = fadd2(v[1], bf2_to_f2(r0.y));
                    v[2] = fadd2(v[2], bf2_to_f2(r0.z)); v[3] = fadd2(v[3], bf2_to_f2(r0.w));
                    v[4] = fadd2(v[4], bf2_to_f2(r1.x)); v[5] = fadd2(v[5], bf2_to_f2(r1.y));
                    v[6] = fadd2(v[6], bf2_to_f2(r1.z)); v[7] = fadd2(v[7], bf2_to_f2(r1.w));
                }
                if constexpr (HEAD) {     // final_conv on the fp32 values: nothing is staged or stored for this tile
                    const uint32_t ha = smem_u32(smem + plan.head_off) + static_cast<uint32_t>(nb) * 4u;
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < p.head_n) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const ulonglong2 ww = lds_128(ha + static_cast<uint32_t>(o * p.n_pad) * 4u + j * 16);
                                hacc[o] = ffma2(v[2 * j], ww.x, hacc[o]);
                                hacc[o] = ffma2(v[2 * j + 1], ww.y, hacc[o]);
                            }
                        }
                    }
                    continue;
                }
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f0, f1;
                    upk2(v[j], f0, f1);
                    w[j] = pack_bf16x2(f0, f1);
                }
                if (want_rn) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = bf16_lo(w[j]), c2 = bf16_hi(w[j]);
                        out_sumsq = fmaf(a, a, fmaf(c2, c2, out_sumsq));
                    }
                }
                sts_128u(s0, w[0], w[1], w[2], w[3]);
                sts_128u(s1, w[4], w[5], w[6], w[7]);
            }
            if (!released) {                  // a warp that read nothing still owes its arrival
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
            }
            if (want_rn) gred_b[part * kTileM + r] = out_sumsq;
            [[maybe_unused]] float hsum[4];
            if constexpr (HEAD) {          // the column parts' partial dots meet in shared memory (kGParts == 2)
#pragma unroll
                for (int o = 0; o < 4; ++o) { float a, b; upk2(hacc[o], a, b); hsum[o] = a + b; }
                if (part == 1)
                    *reinterpret_cast<float4*>(smem + plan.head_off + p.head_n * p.n_pad * 4 + (grp * kTileM + r) * 16) =
                        make_float4(hsum[0], hsum[1], hsum[2], hsum[3]);
            }
            fence_proxy_async();
            if (store_leader) trace_ev(tr, 3 + grp, 4, q, trn);
            {   // one barrier less per tile: the store barrier also publishes "the group's next accumulator is full"
                int n_next, m_next;
                acc_ready = merge_acc && seq_tile(q + GROUPS, n_next, m_next);
                if (acc_ready && store_leader) {
                    const int qn = q + GROUPS;
                    mbar_wait(&bars->acc_full[qn & acc_mask], static_cast<uint32_t>(qn >> acc_shift) & 1u);
                }
            }
            named_bar_sync(bar0 + 1, kGroupThreads);
            if (store_leader) trace_ev(tr, 3 + grp, 5, q, trn);
            if constexpr (HEAD) {
                if (part == 0 && !skip && real_tile) {
                    const float4 o1 = *reinterpret_cast<const float4*>(smem + plan.head_off + p.head_n * p.n_pad * 4 + (grp * kTileM + r) * 16);
                    const float tot[4] = {hsum[0] + o1.x, hsum[1] + o1.y, hsum[2] + o1.z, hsum[3] + o1.w};
                    const int x = tg.x0 + (r & (p.bw - 1)), y = tg.y0 + ((r >> p.bw_shift) & (p.bh - 1));
                    const int b = tg.b0 + (r >> (p.bw_shift + p.bh_shift));
                    if (x < p.W && y < p.H && b < p.B) {
#pragma unroll
                        for (int o = 0; o < 4; ++o)
                            if (o < p.head_n)
                                p.head_out[((static_cast<long long>(b) * p.head_n + o) * p.H + y) * p.W + x] = tot[o] + __ldg(p.head_b + o);
                    }
                }
                // (part 1 rewrites the scratch row for the group's next tile only behind that tile's pre-norm barrier)
            } else if (store_leader) {
                if (!skip) {
                    const int groups = (min(p.block_n, p.N - n0) + 63) >> 6;
                    for (int g = 0; g < groups; ++g) {
                        const int ch = n0 + g * 64;
                        if (p.sy == 2) {     // sub-pixel phase: output viewed as [B, H, (py), W, (px c)]
                            tma_store_5d(&tmOut, buf + g * (kTileM * 128), p.ox * p.ld_out + ch, tg.x0, p.oy, tg.y0, tg.b0);
                        } else {
                            tma_store_5d(&tmOut, buf + g * (kTileM * 128), ch, tg.x0, 0, tg.y0, tg.b0);
                        }
                    }
                }
                bulk_commit_group();         // (possibly empty) group: keeps the wait_group arithmetic uniform
            }
            if (want_rn && part == 0 && !skip && real_tile) {
                float t = gred_b[r];
#pragma unroll
                for (int i = 1; i < kGParts; ++i) t += gred_b[i * kTileM + r];
                p.rnorm_out[tile_pix] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
            }
        }
        if (store_leader) bulk_wait_group<0>();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 3..18)
        // thread = (accumulator row r, column part): 4 parts x 128 rows.  Two passes over TMEM when RMSNorm is on
        // (sum of squares, then normalise); partial sums of the 4 parts meet in shared memory.
        const int ew = warp - 3;
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int part = ew >> 2;               // column part handled by this warp
        const int r = q * 32 + lane;            // accumulator row == tile pixel
        const int bx = r & (p.bw - 1);          // bw, bh are powers of two
        const int by = (r >> p.bw_shift) & (p.bh - 1);
        const int bi = r >> (p.bw_shift + p.bh_shift);
        const bool store_leader = (ew == 0) && (lane == 0);
        const bool ksplit = false;
        const bool need_pix = (p.row_scale != nullptr) || (p.rnorm_out != nullptr) || !p.tma_store || ss_batched;
        const int tiles_xy = p.tiles_x * p.tiles_y;
        int acc = 0;
        uint32_t acc_phase = 0;
        int n_tile, m_tile;
        for (int tq = 0; seq_tile(tq, n_tile, m_tile); ++tq) {   // tq: tile sequence number (q is the TMEM lane quarter)
            int tx, ty, tb;
            if (p.tiles_pow2) {
                tx = m_tile & (p.tiles_x - 1);
                ty = (m_tile >> p.tx_shift) & (p.tiles_y - 1);
                tb = m_tile >> (p.tx_shift + p.ty_shift);
            } else {
                tb = m_tile / tiles_xy;
                const int rem = m_tile - tb * tiles_xy;
                ty = rem / p.tiles_x;
                tx = rem - ty * p.tiles_x;
            }
            const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
            const int n0 = n_tile * p.block_n;
            // Per-pixel addressing is only needed off the fast path (TMA store clips out-of-range rows by itself):
            // pre-norm row scales, the row-norm side output, direct stores, per-sample scale/shift, direct residuals.
            int b = 0, oyy = 0, oxx = 0;
            bool valid = true;
            long long out_pix = 0;
            float rs = 1.0f;
            const float* ssb = nullptr;
            if (need_pix) {
                const int x = x0 + bx, y = y0 + by;
                b = b0 + bi;
                valid = (x < p.W) && (y < p.H) && (b < p.B);
                oyy = y * p.sy + p.oy;
                oxx = x * p.sx + p.ox;
                out_pix = (static_cast<long long>(b) * p.OH + oyy) * p.OW + oxx;
                if (p.row_scale != nullptr && valid) rs = __ldg(p.row_scale + (static_cast<long long>(b) * p.H + y) * p.W + x);
                if (ss_batched) ssb = p.scale_shift + static_cast<long long>(valid ? b : 0) * p.ss_stride;
            }

            const int ncols = min(p.block_n, p.N - n0);
            const int nchunks = (ncols + 15) >> 4;
            const int per = (nchunks + kParts - 1) / kParts;
            const int c_lo = min(nchunks, part * per);
            const int c_hi = min(nchunks, c_lo + per);

            // Residual tile.  With TMA stores the tile is first copied, coalesced and asynchronously (cp.async), into
            // the staging buffer in the staging layout, before waiting for the accumulator: pass 2 then reads its
            // residual from shared memory and overwrites it with the result.  Otherwise each thread prefetches the
            // first two 16-column chunks of its own row into registers.
            const bool res_smem = (p.residual != nullptr) && p.tma_store;
            uint4 rpre[2][2];
            const bool res_vec = (p.residual != nullptr) && !res_smem && valid && ((p.N & 15) == 0);
            const __nv_bfloat16* rrow = p.residual != nullptr ? p.residual + out_pix * p.ld_res + n0 : nullptr;
            if (res_smem) {
                if (store_leader) bulk_wait_group_read<0>();       // previous tile's TMA store has drained the buffer
                named_bar_sync(kBarRes, kEpiThreads);
                const int upr = ncols >> 3;                         // 16-byte units per row
                const int et = threadIdx.x - 96;                    // 0..511
                for (int u = et; u < kTileM * upr; u += kEpiThreads) {
                    const int row = u / upr, cu = u - row * upr;
                    const int rx = x0 + (row & (p.bw - 1));
                    const int ry = y0 + ((row >> p.bw_shift) & (p.bh - 1));
                    const int rb = b0 + (row >> (p.bw_shift + p.bh_shift));
                    uint8_t* dst = staging + (cu >> 3) * (kTileM * 128) + row * 128 + (((cu & 7) ^ (row & 7)) << 4);
                    if (rx < p.W && ry < p.H && rb < p.B) {
                        const long long pix = (static_cast<long long>(rb) * p.OH + (ry * p.sy + p.oy)) * p.OW + (rx * p.sx + p.ox);
                        cp_async_16(dst, p.residual + pix * p.ld_res + n0 + cu * 8);
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                    }
                }
                cp_async_commit();
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (res_vec && c_lo + i < c_hi) {
                    rpre[i][0] = __ldg(reinterpret_cast<const uint4*>(rrow + (c_lo + i) * 16));
                    rpre[i][1] = __ldg(reinterpret_cast<const uint4*>(rrow + (c_lo + i) * 16) + 1);
                }
            }

            if (lane == 0) mbar_wait(&bars->acc_full[acc], acc_phase);   // one polling lane per warp: 512 pollers saturate the smem pipe
            __syncwarp();
            tc_fence_after();
            if ((p.debug & kDbg) & 1) {                // profiling: epilogue does nothing but release the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                continue;
            }
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.acc_stride);

            if (p.norm_g != nullptr) {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                for (int c = c_lo; c < c_hi; ++c) {
                    __syncwarp();
                    uint32_t v[16];
                    tmem_ld16(t_row + c * 16, v);
                    if (ksplit) {
                        uint32_t v2[16];
                        tmem_ld16(t_row + p.block_n + c * 16, v2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    } else {
                        tmem_ld_wait();
                    }
                    const float4* b4 = reinterpret_cast<const float4*>(col_bias + n0 + c * 16);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bb4 = b4[j4];
                        const float f0 = fmaf(__uint_as_float(v[4 * j4 + 0]), rs, bb4.x);
                        const float f1 = fmaf(__uint_as_float(v[4 * j4 + 1]), rs, bb4.y);
                        const float f2 = fmaf(__uint_as_float(v[4 * j4 + 2]), rs, bb4.z);
                        const float f3 = fmaf(__uint_as_float(v[4 * j4 + 3]), rs, bb4.w);
                        s0 = fmaf(f0, f0, s0);
                        s1 = fmaf(f1, f1, s1);
                        s2 = fmaf(f2, f2, s2);
                        s3 = fmaf(f3, f3, s3);          // padded columns have acc == 0 and bias == 0
                    }
                }
                red_a[part * kTileM + r] = (s0 + s1) + (s2 + s3);
            }
            // the previous tile's TMA stores must have finished reading the staging buffer before it is rewritten
            if (p.tma_store && store_leader) bulk_wait_group_read<0>();
            if (res_smem) cp_async_wait_all();     // own residual copies landed; the barrier publishes everyone's
            named_bar_sync(kBarPre, kEpiThreads);
            float rinv = 1.0f;
            if (p.norm_g != nullptr) {
                float tot = (red_a[r] + red_a[kTileM + r]) + (red_a[2 * kTileM + r] + red_a[3 * kTileM + r]);
                if (p.pair_n) {      // the other half of the row lives in the peer CTA: swap the halves' sums of squares
                    float* xch = reinterpret_cast<float*>(smem + plan.xch_off);
                    const int par = tq & 1;
                    if (part == 0) {
                        const uint32_t peer = static_cast<uint32_t>(cl_rank ^ 1);
                        st_cluster_f32(cluster_map(smem_u32(xch + par * kTileM + r), peer), tot);
                        mbar_arrive_cluster(cluster_map(smem_u32(&bars->xch_full[par]), peer));
                    }
                    if (lane == 0) mbar_wait_cluster(&bars->xch_full[par], static_cast<uint32_t>(tq >> 1) & 1u);
                    __syncwarp();
                    const float mine = tot, theirs = xch[par * kTileM + r];
                    tot = cl_rank == 0 ? mine + theirs : theirs + mine;      // same operand order in both CTAs: identical bits
                }
                rinv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
            }

            float out_sumsq = 0.0f;
            for (int c = c_lo; c < c_hi; ++c) {
                __syncwarp();
                uint32_t v[16];
                tmem_ld16(t_row + c * 16, v);
                if (ksplit) {
                    uint32_t v2[16];
                    tmem_ld16(t_row + p.block_n + c * 16, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                } else {
                    tmem_ld_wait();
                }
                if (c == c_hi - 1) {
                    // last TMEM read of this accumulator stage: hand it back to the MMA warp before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);     // one arrival per epilogue warp
                }
                float f[16];
                const int nb = n0 + c * 16;
                {
                    const float4* b4 = reinterpret_cast<const float4*>(col_bias + nb);
                    const float4* m4 = reinterpret_cast<const float4*>(col_mul + nb);
                    const float4* a4 = reinterpret_cast<const float4*>(col_add + nb);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bb4 = b4[j4];
                        f[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), rs, bb4.x);
                        f[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), rs, bb4.y);
                        f[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), rs, bb4.z);
                        f[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), rs, bb4.w);
                        if (affine) {
                            const float4 mm = m4[j4], aa = a4[j4];
                            f[4 * j4 + 0] = fmaf(f[4 * j4 + 0] * rinv, mm.x, aa.x);
                            f[4 * j4 + 1] = fmaf(f[4 * j4 + 1] * rinv, mm.y, aa.y);
                            f[4 * j4 + 2] = fmaf(f[4 * j4 + 2] * rinv, mm.z, aa.z);
                            f[4 * j4 + 3] = fmaf(f[4 * j4 + 3] * rinv, mm.w, aa.w);
                        }
                    }
                }
                if (ssb != nullptr) {          // per-sample time embedding (generic forward, not the sampling loop)
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (nb + j < p.N) f[j] = fmaf(f[j], __ldg(ssb + nb + j) + 1.0f, __ldg(ssb + p.N + nb + j));
                }
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = silu_f(f[j]);
                }
                if (SPLITK && p.partial != nullptr) {      // split-K: raw fp32 sums of this K range (bias, norm, ... happen in the consumer)
                    if (valid && nb < p.N) {
                        int lo_u, hi_u, ks;
                        stage_range(tq, lo_u, hi_u, ks);
                        float4* o = reinterpret_cast<float4*>(p.partial + ks * p.partial_stride + out_pix * p.N + nb);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            if (nb + 4 * j4 < p.N) o[j4] = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
                    }
                    continue;
                }
                if (p.out_f32_nchw) {
                    if (valid) {
                        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = nb + j;
                            if (n < p.N) o[((static_cast<long long>(b) * p.N + n) * p.OH + oyy) * p.OW + oxx] = f[j];
                        }
                    }
                    continue;
                }
                if (res_smem) {
                    const int cl = c * 16;
                    const uint8_t* rowp = staging + (cl >> 6) * (kTileM * 128) + r * 128;
                    const int u = (cl & 63) >> 3;
                    const uint4 r0 = *reinterpret_cast<const uint4*>(rowp + (((u) ^ (r & 7)) << 4));
                    const uint4 r1 = *reinterpret_cast<const uint4*>(rowp + (((u + 1) ^ (r & 7)) << 4));
                    f[0] += bf16_lo(r0.x); f[1] += bf16_hi(r0.x); f[2] += bf16_lo(r0.y); f[3] += bf16_hi(r0.y);
                    f[4] += bf16_lo(r0.z); f[5] += bf16_hi(r0.z); f[6] += bf16_lo(r0.w); f[7] += bf16_hi(r0.w);
                    f[8] += bf16_lo(r1.x); f[9] += bf16_hi(r1.x); f[10] += bf16_lo(r1.y); f[11] += bf16_hi(r1.y);
                    f[12] += bf16_lo(r1.z); f[13] += bf16_hi(r1.z); f[14] += bf16_lo(r1.w); f[15] += bf16_hi(r1.w);
                } else if (p.residual != nullptr && valid) {
                    if (res_vec) {
                        uint4 r0, r1;
                        const int i = c - c_lo;
                        if (i == 0) { r0 = rpre[0][0]; r1 = rpre[0][1]; }
                        else if (i == 1) { r0 = rpre[1][0]; r1 = rpre[1][1]; }
                        else {
                            r0 = __ldg(reinterpret_cast<const uint4*>(rrow + c * 16));
                            r1 = __ldg(reinterpret_cast<const uint4*>(rrow + c * 16) + 1);
                        }
                        f[0] += bf16_lo(r0.x); f[1] += bf16_hi(r0.x); f[2] += bf16_lo(r0.y); f[3] += bf16_hi(r0.y);
                        f[4] += bf16_lo(r0.z); f[5] += bf16_hi(r0.z); f[6] += bf16_lo(r0.w); f[7] += bf16_hi(r0.w);
                        f[8] += bf16_lo(r1.x); f[9] += bf16_hi(r1.x); f[10] += bf16_lo(r1.y); f[11] += bf16_hi(r1.y);
                        f[12] += bf16_lo(r1.z); f[13] += bf16_hi(r1.z); f[14] += bf16_lo(r1.w); f[15] += bf16_hi(r1.w);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < p.N) f[j] += __bfloat162float(rrow[c * 16 + j]);
                    }
                }
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                if (p.rnorm_out != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = bf16_lo(w[j]), c2 = bf16_hi(w[j]);
                        out_sumsq = fmaf(a, a, fmaf(c2, c2, out_sumsq));   // padded columns are exactly 0
                    }
                }
                if (p.tma_store) {
                    // staging: [group of 64 channels][128 rows][128 B], 16-byte units XOR-swizzled by (row & 7)
                    const int cl = c * 16;                       // column inside this N tile
                    uint8_t* rowp = staging + (cl >> 6) * (kTileM * 128) + r * 128;
                    const int u = (cl & 63) >> 3;                // 16-byte unit inside the 128-byte row
                    *reinterpret_cast<uint4*>(rowp + (((u) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(rowp + (((u + 1) ^ (r & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                } else if (valid) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_pix * p.ld_out + nb;
                    if (nb + 16 <= p.N) {
                        reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (nb + 2 * j < p.N) o[2 * j] = __ushort_as_bfloat16(static_cast<unsigned short>(w[j] & 0xFFFFu));
                            if (nb + 2 * j + 1 < p.N) o[2 * j + 1] = __ushort_as_bfloat16(static_cast<unsigned short>(w[j] >> 16));
                        }
                    }
                }
            }
            if (c_lo >= c_hi) {               // a warp with no columns still owes its arrival
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
            }
            if (p.rnorm_out != nullptr) red_b[part * kTileM + r] = out_sumsq;
            if (p.tma_store) fence_proxy_async();
            named_bar_sync(kBarPost, kEpiThreads);
            if (p.tma_store && store_leader) {
                const int groups = (ncols + 63) >> 6;
                for (int g = 0; g < groups; ++g) {
                    const int ch = n0 + g * 64;
                    if (p.sy == 2) {     // sub-pixel phase: output viewed as [B, H, (py), W, (px c)]
                        tma_store_5d(&tmOut, staging + g * (kTileM * 128), p.ox * p.ld_out + ch, x0, p.oy, y0, b0);
                    } else {
                        tma_store_5d(&tmOut, staging + g * (kTileM * 128), ch, x0, 0, y0, b0);
                    }
                }
                bulk_commit_group();
            }
            if (p.rnorm_out != nullptr && part == 0 && valid)
                p.rnorm_out[out_pix] =
                    1.0f / fmaxf(sqrtf((red_b[r] + red_b[kTileM + r]) + (red_b[2 * kTileM + r] + red_b[3 * kTileM + r])), 1e-12f);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (p.tma_store && store_leader) bulk_wait_group<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster == 2) cluster_sync_all();      // no CTA may exit while its peer can still multicast into it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
}

}  // namespace

int conv_smem_plan(const ConvParams& p, int* num_stages) {
    const int slack = p.tight_smem ? 0 : 1024;    // alignment slack
    const int budget = 227 * 1024 - slack;
    int stages = 8;
    while (stages > 0 && make_plan(p, stages).total > budget) --stages;
    *num_stages = stages;
    return make_plan(p, stages > 0 ? stages : 1).total + slack;
}

int conv_trace_read(long long* host, int cap) {
    static long long tmp[kTraceRoles * kTraceCap * 2];
    cudaMemcpyFromSymbol(tmp, g_trace, sizeof(tmp));
    int n = 0;
    for (int i = 0; i < kTraceRoles * kTraceCap && n < cap; ++i)
        if (tmp[2 * i + 1] != 0) { host[2 * n] = tmp[2 * i]; host[2 * n + 1] = tmp[2 * i + 1]; ++n; }
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, g_trace);
    cudaMemset(sym, 0, sizeof(tmp));
    return n;
}

int conv_prepare_attributes() {
    int r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, false, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, false, 0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 0, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (r == 0) r = static_cast<int>(cudaFuncSetAttribute(conv_tc_kernel<16, true, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return r;
}

void launch_conv(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW, const CUtensorMap& tmOut,
                 const CUtensorMap& tmRes, const CUtensorMap& tmR1, const ConvParams& p, int num_sms, cudaStream_t stream, bool pdl) {
    int stages = 0;
    const int smem = conv_smem_plan(p, &stages);
    int grid;
    if (p.cluster == 2) {
        const int cluster_tiles = p.pair_n ? p.m_tiles : p.pairs * p.n_tiles;
        const int n_clusters = cluster_tiles < num_sms / 2 ? cluster_tiles : num_sms / 2;
        grid = 2 * n_clusters;
    } else {
        grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(96 + 32 * 16);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster == 2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // programmatic dependent launch, see ptx.cuh
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    if (p.epi_groups == 4 && p.fold == 2) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 2, 4>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.epi_groups == 4) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 0, 4>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.fold == 3) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 3, 2>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.fold == 2) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 2, 2>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.head_n) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 0, 2, false, true>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.ksplit > 1) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, false, 0, 2, true>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else if (p.fast_epilogue) {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true, 0, 2>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    } else {
        cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, false, 0, 2>, tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p);
    }
}

}  // namespace ddm
