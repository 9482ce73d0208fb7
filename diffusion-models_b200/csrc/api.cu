// C ABI of libddm_b200.so (declared in include/ddm_b200.h).  Host-side only: argument validation, TMA descriptor
// encoding and kernel launches on the caller's stream.  No allocation, no synchronisation, no CPU fallback.
#include "../../include/ddm_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "conv_tc.cuh"
#include "kernels.cuh"

namespace ddm {
void launch_linattn_fused(const CUtensorMap& tmX, const CUtensorMap& tmY, const CUtensorMap& tmWqkv, const CUtensorMap& tmWout,
                          const float* bias_out, const float* g_out, const float* mem_kv, const float* k_shift, int B, int n, int C,
                          int n_mem, int num_sms, int trace, bool pdl, cudaStream_t s);
int linattn_trace_read(long long* host, int cap);
int attention_tc_prepare_attributes();
bool attention_tc_supported(int d, int n_mem);
void launch_attention_tc(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const float* mem_k, const float* mem_v,
                         int n_mem, void* out, int B, int nq, int nk, int heads, int d, cudaStream_t s);
}

namespace {

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_num_sms = 0;
bool g_ready = false;
int g_conv_debug = 0;
int g_laf_trace = 0;
bool g_tc_attention = true;
bool g_pdl = true;             // programmatic dependent launch of the big kernels (DDM_NO_PDL=1 disables, for A/B timing)
std::atomic<long long> g_launches{0};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ddm_conv2d plans (tile / slab / epilogue configuration + the five encoded tensor maps) are pure functions of the argument
// struct: callers that keep their ddm_conv_args alive and unchanged (the engine does, one per layer) get them from this
// cache instead of paying five cuTensorMapEncodeTiled driver calls per launch (~100 launches per eager U-Net forward).
// Keyed by the struct's address, validated by comparing its bytes; bounded.
struct ConvPlan {
    ddm_conv_args args;
    ddm::ConvParams p;
    CUtensorMap tmA0, tmA1, tmW, tmOut, tmRes, tmR1;
};
std::unordered_map<const ddm_conv_args*, ConvPlan> g_plans;
std::mutex g_plans_mu;
constexpr size_t kMaxPlans = 8192;
inline int finish(int launches) {
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    return static_cast<int>(cudaPeekAtLastError());
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds.  dims/strides innermost first; strides in elements.
int encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                    const unsigned long long* strides_elems, const unsigned* box) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 1; i < rank; ++i) {
        gstr[i - 1] = strides_elems[i] * 2ull;
        if (gstr[i - 1] % 16ull != 0) return DDM_E_ALIGNMENT;
    }
    if (!aligned16(base)) return DDM_E_ALIGNMENT;
    const CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                                gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : DDM_E_DRIVER;
}

}  // namespace

extern "C" {

int ddm_abi_version(void) { return DDM_ABI_VERSION; }

long long ddm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* ddm_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case DDM_E_NOT_INITIALISED: return "ddm_init has not been called (or failed)";
        case DDM_E_BAD_ARGUMENT: return "bad argument";
        case DDM_E_UNSUPPORTED: return "unsupported shape or option";
        case DDM_E_ALIGNMENT: return "pointer or stride not 16-byte aligned";
        case DDM_E_DRIVER: return "CUDA driver call failed (cuTensorMapEncodeTiled / entry point)";
        case DDM_E_WRONG_ARCH: return "device is not compute capability 10.x (B200, sm_100a)";
        default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown ddm error";
    }
}

static int init_on_current_device(int device);

int ddm_init(int device) {
    // the attribute calls below act on the current device: switch to `device` for their duration only
    int prev = -1;
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int r = init_on_current_device(device);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return r;
}

static int init_on_current_device(int device) {
    cudaError_t e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (prop.major != 10) return DDM_E_WRONG_ARCH;
    g_num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return DDM_E_DRIVER;
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    int r = ddm::conv_prepare_attributes();
    if (r != 0) return r;
    r = ddm::stem_prepare_attributes();
    if (r != 0) return r;
    r = ddm::stem_tc_prepare_attributes();
    if (r == 0) r = ddm::stem_umma_prepare_attributes();
    if (r != 0) return r;
    r = ddm::linattn_fused_prepare_attributes();
    if (r != 0) return r;
    r = ddm::attention_tc_prepare_attributes();
    if (r != 0) return r;
    if (std::getenv("DDM_NO_TC_ATTENTION")) g_tc_attention = false;
    if (std::getenv("DDM_NO_PDL")) g_pdl = false;      // A/B timing against the CUDA-core kernel
    if (const char* t = std::getenv("DDM_LAF_TRACE")) g_laf_trace = std::atoi(t);       // scripts/laf_trace.py
    if (const char* dbg = std::getenv("DDM_CONV_DEBUG")) g_conv_debug = std::atoi(dbg);   // bottleneck bisection, see conv_tc.cuh
    g_ready = true;
    return 0;
}

int ddm_conv2d(const ddm_conv_args* a, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    const bool splitk = a != nullptr && a->ksplit > 1;
    const bool head = a != nullptr && a->head_out != nullptr;
    if (a == nullptr || a->src0 == nullptr || a->weight == nullptr || (a->out == nullptr && !splitk && !head)) return DDM_E_BAD_ARGUMENT;
    if (head) {
        if (a->head_w == nullptr || a->head_b == nullptr || splitk || a->out_f32_nchw || a->rnorm_out != nullptr || a->view != 0 ||
            a->sy != 1 || a->sx != 1 || a->OH != a->H || a->OW != a->W)
            return DDM_E_BAD_ARGUMENT;
        if (!ddm_conv2d_head_supported(a->N, a->head_n, a->H, a->W) || a->N_pad != a->N || a->norm_g == nullptr ||
            (a->residual == nullptr && a->rsrc0 == nullptr))       // (plain convs take the dx-folded kernels, which have no head)
            return DDM_E_UNSUPPORTED;
    }
    if (splitk && (a->partial_out == nullptr || !aligned16(a->partial_out) || a->bias != nullptr || a->row_scale != nullptr ||
                   a->norm_g != nullptr || a->scale_shift != nullptr || a->act != 0 || a->residual != nullptr || a->rnorm_out != nullptr ||
                   a->rsrc0 != nullptr || a->view != 0 || a->out_f32_nchw || a->sy != 1 || a->sx != 1 || a->OH != a->H || a->OW != a->W ||
                   (a->N % 4) != 0 || a->ksplit > 64))
        return DDM_E_BAD_ARGUMENT;
    if (a->ntaps < 1 || a->ntaps > DDM_MAX_TAPS || a->B < 1 || a->H < 1 || a->W < 1 || a->N < 1) return DDM_E_BAD_ARGUMENT;
    if (a->C0 < 8 || (a->C0 % 8) != 0 || (a->src1 != nullptr && (a->C1 < 8 || (a->C1 % 8) != 0))) return DDM_E_UNSUPPORTED;
    if ((a->ld0 % 8) != 0 || (a->src1 != nullptr && (a->ld1 % 8) != 0)) return DDM_E_ALIGNMENT;
    if (a->view != 0 && a->view != 1) return DDM_E_BAD_ARGUMENT;
    if (a->view == 1 && a->src1 != nullptr) return DDM_E_UNSUPPORTED;
    if ((a->N_pad % 16) != 0 || a->N_pad < a->N || (a->K_pad % 64) != 0) return DDM_E_BAD_ARGUMENT;
    if (!splitk && !head && !a->out_f32_nchw && ((a->ld_out % 8) != 0 || !aligned16(a->out))) return DDM_E_ALIGNMENT;
    if (a->residual != nullptr && ((a->ld_res % 8) != 0 || !aligned16(a->residual))) return DDM_E_ALIGNMENT;
    const bool shortcut = a->rsrc0 != nullptr;
    if (shortcut) {
        if (a->residual != nullptr || a->rC0 < 64 || (a->rC0 % 64) != 0 || (a->rld0 % 8) != 0 || a->view != 0) return DDM_E_BAD_ARGUMENT;
        if (a->rsrc1 != nullptr ? (a->rC1 < 64 || (a->rC1 % 64) != 0 || (a->rld1 % 8) != 0) : a->rC1 != 0) return DDM_E_BAD_ARGUMENT;
        if (!ddm_conv2d_shortcut_supported(a->N, a->C0 + a->C1, a->rC0, a->rC1, a->H, a->W) || a->ntaps != 9 || a->norm_g == nullptr ||
            a->rnorm_out != nullptr || a->out_f32_nchw || a->OH != a->H || a->OW != a->W || a->sy != 1 || a->sx != 1 || a->N_pad != a->N)
            return DDM_E_UNSUPPORTED;
    }
    {
        std::lock_guard<std::mutex> lock(g_plans_mu);
        auto it = g_plans.find(a);
        if (it != g_plans.end() && std::memcmp(&it->second.args, a, sizeof(*a)) == 0) {
            const ConvPlan& c = it->second;
            ddm::launch_conv(c.tmA0, c.tmA1, c.tmW, c.tmOut, c.tmRes, c.tmR1, c.p, g_num_sms, as_stream(stream), g_pdl);
            return finish(1);
        }
    }

    ddm::ConvParams p;
    std::memset(&p, 0, sizeof(p));
    p.B = a->B; p.H = a->H; p.W = a->W;
    // tile box: 128 pixels = bw x bh x bb (powers of two), x fastest
    // at most 32 wide: the A slab of a tile is (bh + 2) x bw pixels, so squarer tiles carry less halo through L2
    // (128x1 tiles re-read every input row 3 times, 32x4 tiles 1.5 times); DDM_CONV_DEBUG & 262144 restores the wide tiles
    p.bw = (g_conv_debug & 262144) ? (a->W >= 128 ? 128 : pow2_ceil(a->W)) : (a->W >= 32 ? 32 : pow2_ceil(a->W));
    p.bh = pow2_ceil(a->H); if (p.bh > 128 / p.bw) p.bh = 128 / p.bw;
    p.bb = 128 / (p.bw * p.bh);
    p.tiles_x = (a->W + p.bw - 1) / p.bw;
    p.tiles_y = (a->H + p.bh - 1) / p.bh;
    const int tiles_b = (a->B + p.bb - 1) / p.bb;
    p.m_tiles = p.tiles_x * p.tiles_y * tiles_b;
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    p.bw_shift = ilog2(p.bw); p.bh_shift = ilog2(p.bh);
    p.tx_shift = ilog2(p.tiles_x); p.ty_shift = ilog2(p.tiles_y);
    p.tiles_pow2 = ((1 << p.tx_shift) == p.tiles_x && (1 << p.ty_shift) == p.tiles_y) ? 1 : 0;
    if (a->N_pad > ddm::kMaxNPad) return DDM_E_UNSUPPORTED;
    p.n_tiles = (a->N_pad + 255) / 256;
    p.block_n = ((a->N_pad + p.n_tiles - 1) / p.n_tiles + 15) / 16 * 16;   // e.g. N=384 -> 2 tiles of 192
    // C_out = 384 (to_qkv) with one 64-channel K chunk and no fused norm: three tiles of 128 instead of two of 192, so that
    // four accumulator stages fit in TMEM and the four-group lean epilogue (single pass here) applies.  Measured at
    // B = 1024: 208 -> 192 us at 32x32, 58.5 -> 55 us at 16x16; with K = 128 the 2 x 192 split stays faster.
    // DDM_CONV_DEBUG & 16777216 keeps 2 x 192.
    const bool qkv_three_tiles = a->N_pad == 384 && a->K_pad == 64 && a->norm_g == nullptr && a->rnorm_out == nullptr &&
                                 a->residual == nullptr && !(g_conv_debug & 16777216);
    if (qkv_three_tiles) { p.n_tiles = 3; p.block_n = 128; }
    // Small M (the 4x4 / 8x8 levels at small per-GPU batches): with 256-wide N tiles a 512-channel layer at M = 2048 is 32 tiles
    // for 148 SMs.  Narrower N tiles multiply the tile count at the same total weight traffic (each tile streams its own
    // slice of W; only the small A operand is re-read, from L2): halve block_n while that still leaves most SMs idle.
    // Not with a fused row norm (one tile must own the whole output row).  DDM_CONV_DEBUG & 33554432 disables it.
    if (a->norm_g == nullptr && a->rnorm_out == nullptr && !qkv_three_tiles && !splitk && !(g_conv_debug & 33554432)) {
        while (p.m_tiles * p.n_tiles * 2 <= g_num_sms && p.block_n >= 128 && (p.block_n % 128) == 0 &&      // tiles stay 64-channel
               (a->N_pad % (p.block_n / 2)) == 0) {                                                         // groups (TMA store)
            p.block_n /= 2;
            p.n_tiles = a->N_pad / p.block_n;
        }
    }
    p.ksplit = splitk ? a->ksplit : 1;
    p.total_tiles = p.m_tiles * p.n_tiles * p.ksplit;
    p.N = a->N;
    p.rnorm_out = a->rnorm_out;       // (the shared-memory plan depends on it)
    // a normalised row wider than one accumulator: its two N tiles in a CTA pair (conv_tc.cuh: pair_n)
    const bool pair_norm = a->norm_g != nullptr && p.n_tiles == 2 && ddm_conv2d_row_norm_supported(a->N) && a->N == a->N_pad &&
                           a->rnorm_out == nullptr && !a->out_f32_nchw && !shortcut && !splitk && g_num_sms >= 2;
    if (a->norm_g != nullptr && p.n_tiles != 1 && !pair_norm) return DDM_E_UNSUPPORTED;
    if (a->rnorm_out != nullptr && (p.n_tiles != 1 || a->out_f32_nchw)) return DDM_E_UNSUPPORTED;
    // group taps into slabs (same dx and p, consecutive dy) when the dy shift is expressible as an aligned row offset
    auto set_slabs = [&](bool want_group) {
        bool can_group = want_group && (p.bb == 1) && (p.bw % 8 == 0);
        int gdx[DDM_MAX_TAPS], gp[DDM_MAX_TAPS], gcnt[DDM_MAX_TAPS], gdy[DDM_MAX_TAPS][DDM_MAX_TAPS], gtap[DDM_MAX_TAPS][DDM_MAX_TAPS];
        int ng = 0;
        for (int t = 0; t < a->ntaps && can_group; ++t) {
            int g = 0;
            while (g < ng && !(gdx[g] == a->tap_dx[t] && gp[g] == a->tap_p[t])) ++g;
            if (g == ng) { gdx[g] = a->tap_dx[t]; gp[g] = a->tap_p[t]; gcnt[g] = 0; ++ng; }
            gdy[g][gcnt[g]] = a->tap_dy[t]; gtap[g][gcnt[g]] = t; ++gcnt[g];
        }
        for (int g = 0; g < ng && can_group; ++g) {
            if (gcnt[g] != gcnt[0] || gcnt[g] > 3) can_group = false;
            for (int j = 0; j < gcnt[g] && can_group; ++j) {     // insertion sort by dy, then require consecutive dy
                for (int i = j; i > 0 && gdy[g][i] < gdy[g][i - 1]; --i) {
                    int tdy = gdy[g][i]; gdy[g][i] = gdy[g][i - 1]; gdy[g][i - 1] = tdy;
                    int tt = gtap[g][i]; gtap[g][i] = gtap[g][i - 1]; gtap[g][i - 1] = tt;
                }
            }
            for (int j = 1; j < gcnt[g] && can_group; ++j) if (gdy[g][j] != gdy[g][j - 1] + 1) can_group = false;
        }
        if (can_group && ng > 0 && gcnt[0] > 1) {
            p.n_slabs = ng; p.n_dy = gcnt[0];
            for (int g = 0; g < ng; ++g) {
                p.slab_dx[g] = gdx[g]; p.slab_p[g] = gp[g]; p.slab_dy0[g] = gdy[g][0];
                for (int j = 0; j < p.n_dy; ++j) p.slab_tap[g][j] = gtap[g][j];
            }
        } else {
            p.n_slabs = a->ntaps; p.n_dy = 1;
            for (int t = 0; t < a->ntaps; ++t) {
                p.slab_dx[t] = a->tap_dx[t]; p.slab_p[t] = a->tap_p[t]; p.slab_dy0[t] = a->tap_dy[t]; p.slab_tap[t][0] = t;
            }
        }
        p.a_rows = p.bw * (p.bh + p.n_dy - 1) * p.bb;
        return p.n_dy > 1;
    };
    const int ceff0 = a->view == 1 ? 2 * a->C0 : a->C0;
    p.chunks0 = (ceff0 + 63) / 64;
    p.chunks1 = a->src1 != nullptr ? (a->C1 + 63) / 64 : 0;
    p.res_chunks0 = shortcut ? a->rC0 / 64 : 0;
    p.res_chunks = shortcut ? (a->rC0 + a->rC1) / 64 : 0;
    p.rbias = a->rbias;
    if ((a->ntaps * (p.chunks0 + p.chunks1) + p.res_chunks) * 64 != a->K_pad) return DDM_E_BAD_ARGUMENT;
    p.k_chunks = a->ntaps * (p.chunks0 + p.chunks1);
    p.n_pad = a->N_pad;
    if (p.n_pad > ddm::kMaxNPad) return DDM_E_UNSUPPORTED;
    // bf16 tiles whose channel count is a multiple of 64 leave through smem staging + TMA stores
    const bool strided_out = (a->sy != 1 || a->sx != 1);
    if (strided_out && !(a->sy == 2 && a->sx == 2 && a->OH == 2 * a->H && a->OW == 2 * a->W)) return DDM_E_UNSUPPORTED;
    p.head_n = head ? a->head_n : 0; p.head_w = a->head_w; p.head_b = a->head_b; p.head_out = a->head_out;
    p.tma_store = (!splitk && !a->out_f32_nchw && (a->N % 64) == 0 && (!strided_out || a->ld_out == a->N) && a->OH >= a->H * a->sy &&
                   a->OW >= a->W * a->sx) ? 1 : 0;
    {   // lean epilogue kernel: staged TMA store, batch-shared scale/shift, and full tiles wherever a per-pixel side
        // input/output (row_scale, rnorm_out) is addressed by tile offset
        const bool full_tiles = (a->W % p.bw == 0) && (a->H % p.bh == 0) && (a->B % p.bb == 0);
        const bool side = (a->row_scale != nullptr) || (a->rnorm_out != nullptr);
        const bool batched_ss = (a->scale_shift != nullptr) && (a->ss_stride != 0);
        p.fast_epilogue = (p.tma_store && !batched_ss && (!side || (full_tiles && !strided_out)) && !pair_norm && !(g_conv_debug & 8)) ? 1 : 0;
        p.staging_bufs = 1;
    }
    // shared-memory configuration: prefer A-slab reuse and resident weights, as long as >= 3 pipeline stages remain
    {
        bool done = false;
        // (split-K: one pipeline stage per 64-channel K chunk, so that ddm_conv2d_suggest_ksplit's K_pad / 64 is the stage
        // count, and streamed weights -- a CTA only touches its own K range)
        for (int grouped = splitk ? 0 : 1; grouped >= 0 && !done; --grouped) {
            if (grouped && !set_slabs(true)) continue;
            if (!grouped) set_slabs(false);
            for (int resident = (p.n_tiles == 1 && !splitk ? 1 : 0); resident >= 0 && !done; --resident) {
                p.b_resident = resident;
                ddm::conv_smem_plan(p, &p.num_stages);
                if (p.num_stages >= (grouped || resident ? 3 : 2)) done = true;
            }
        }
        if (!done) return DDM_E_UNSUPPORTED;
        if (p.fast_epilogue) {          // the lean kernel runs two epilogue groups, each with its own staging buffer
            int st = 0;
            p.staging_bufs = 2;
            ddm::conv_smem_plan(p, &st);
            if (st >= 2) p.num_stages = st; else { p.staging_bufs = 1; p.fast_epilogue = 0; }
        }
    }
    if (shortcut && !(p.fast_epilogue && p.n_tiles == 1 && (p.block_n == 64 || p.block_n == 128))) return DDM_E_UNSUPPORTED;
    if (head && !(p.fast_epilogue && p.fold == 0 && p.n_tiles == 1 && (p.block_n == 64 || p.block_n == 128))) return DDM_E_UNSUPPORTED;
    p.tmem_cols = pow2_ceil((shortcut ? 4 : 2) * p.block_n); if (p.tmem_cols < 32) p.tmem_cols = 32;
    p.acc_stride = p.tmem_cols / 2;
    p.acc_stages = 2;
    // dx-folded mode (conv_tc.cuh): plain 3x3, C_out padded to 64, tiles spanning whole image rows inside one warp.
    // It cuts the L2 -> shared-memory traffic of the A slabs (the mainloop bound of these layers) at the price of a
    // heavier epilogue, so it is used where the epilogue has slack: no residual input.  DDM_CONV_DEBUG & 512 disables it.
    if (p.fast_epilogue && !(g_conv_debug & 512) && a->residual == nullptr && !shortcut && a->ntaps == 9 && a->view == 0 && p.n_tiles == 1 && p.block_n == 64 &&
        a->N_pad == 64 && p.bw == a->W && p.bw >= 8 && p.bw <= 32 && p.bb == 1 && !strided_out) {
        bool canonical = true;
        unsigned seen = 0;
        for (int t = 0; t < 9; ++t) {
            const int dy = a->tap_dy[t], dx = a->tap_dx[t];
            if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || a->tap_p[t] != 0) { canonical = false; break; }
            seen |= 1u << ((dy + 1) * 3 + (dx + 1));
        }
        if (canonical && seen == 0x1FFu) {
            ddm::ConvParams f = p;
            // fold = 3: one slab, all three dx groups in the accumulator (least L2 traffic, but 3x the TMEM -> register
            // traffic and two shuffles per value in the epilogue).  fold = 2 (default): slabs dx = 0 and dx = +1, two
            // groups, one shuffle: L2, tensor pipe and epilogue come out balanced.  DDM_CONV_DEBUG & 1024 picks 3.
            f.fold = (g_conv_debug & 1024) ? 3 : 2;
            f.n_slabs = f.fold == 3 ? 1 : 2; f.n_dy = 3;
            f.slab_dx[0] = 0; f.slab_p[0] = 0; f.slab_dy0[0] = -1;
            f.slab_dx[1] = 1; f.slab_p[1] = 0; f.slab_dy0[1] = -1;
            for (int t = 0; t < 9; ++t) { f.fold_dyi[t] = a->tap_dy[t] + 1; f.fold_dxi[t] = a->tap_dx[t] + 1; }
            f.a_rows = f.bw * (f.bh + 2);
            f.b_resident = 1;
            f.staging_bufs = 2;
            f.tmem_cols = f.fold == 3 ? 512 : 256; f.acc_stride = f.tmem_cols / 2;
            int st = 0;
            ddm::conv_smem_plan(f, &st);
            if (st < 2 && !(g_conv_debug & 67108864)) {
                // 128 -> 64 (two 64-channel chunks per tap: 144 KB of weights): two 24 KB slab stages and both staging buffers
                // fit only without the alignment slack.  Resident weights are worth it: streamed, every M tile pulls the whole
                // matrix through L2 again (288 KB per tile against 48 KB).  DDM_CONV_DEBUG & 67108864 keeps the streamed plan.
                // Only two stages: the one-slab fold (12 N = 192 MMAs = 1.15 k cycles per stage) is the variant that covers a
                // slab's L2 latency with the other stage's MMAs (measured at B = 1024: streamed 179 us, resident fold 2
                // 187 us, resident fold 3 140 us).
                f.tight_smem = 1;
                f.fold = 3; f.n_slabs = 1; f.tmem_cols = 512; f.acc_stride = 256;
                ddm::conv_smem_plan(f, &st);
            }
            if (st >= 2) { f.num_stages = st; p = f; }
        }
    }
    // Lean epilogue with four groups of four warps (thread = pixel x all columns; conv_tc.cu) where four accumulator stages
    // fit: costs two more staging buffers and a second TMEM pass.  Measured (B = 1024, 32x32): -11 % on the 1x1 conv +
    // RMSNorm + residual of the attention output (to_out), +18 % on 3x3 64->64, +35 % on C_out = 128 -- so it is used
    // for the first kind only.  DDM_CONV_DEBUG & 4194304 forces it wherever it fits, & 8388608 disables it.
    p.epi_groups = 2;
    // (the head lives in the two-group lean kernel)
    const bool four_groups_pays = !head && ((a->ntaps == 1 && a->norm_g != nullptr && a->residual != nullptr && p.block_n == 64) || qkv_three_tiles);
    if (p.fast_epilogue && !head && ((g_conv_debug & 4194304) || (four_groups_pays && !(g_conv_debug & 8388608))) && !(g_conv_debug & 2048) &&
        p.fold != 3) {
        const int cols = (p.fold ? p.fold : (shortcut ? 2 : 1)) * p.block_n;
        const int stride = pow2_ceil(cols) < 32 ? 32 : pow2_ceil(cols);
        if (4 * stride <= 512) {
            ddm::ConvParams f = p;
            f.staging_bufs = 4;
            int st = 0;
            ddm::conv_smem_plan(f, &st);
            if (st >= 2) { f.num_stages = st; f.epi_groups = 4; p = f; }
        }
    }
    // Two MMA issuer threads (conv_tc.cu).  Mode 2, alternate tiles, each thread with its own half of the smem ring, when each
    // half holds a whole tile (both threads' tiles in flight at once): no ordering needed between the threads.  Otherwise mode 1, alternate stages of
    // the same tile in token order.  DDM_CONV_DEBUG & 32: single issuer; & 16384: always mode 1; & 32768: always mode 2.
    {
        const int stages_per_tile = p.n_slabs * (p.chunks0 + p.chunks1) + p.res_chunks;
        p.issue_mode = (g_conv_debug & 32) ? 0 : ((g_conv_debug & 16384) ? 1 : ((g_conv_debug & 32768) ? 2 : ((2 * stages_per_tile <= p.num_stages) ? 2 : 1)));
        if (splitk) {           // stage ranges differ between the tiles of a CTA: the issuers alternate stages (mode 1), never tiles
            p.ks_per = (stages_per_tile + p.ksplit - 1) / p.ksplit;
            if ((p.ksplit - 1) * p.ks_per >= stages_per_tile) return DDM_E_BAD_ARGUMENT;      // an empty range
            if (p.issue_mode == 2) p.issue_mode = 1;
            p.partial = a->partial_out;
            p.partial_stride = static_cast<long long>(a->B) * a->H * a->W * a->N;
        }
    }
    if (p.fast_epilogue && !(g_conv_debug & 2048)) {      // four accumulator stages where TMEM has room (DDM_CONV_DEBUG & 2048: two)
        const int cols = (p.fold ? p.fold : (shortcut ? 2 : 1)) * p.block_n;
        const int stride = pow2_ceil(cols) < 32 ? 32 : pow2_ceil(cols);
        if (4 * stride <= 512) { p.acc_stages = 4; p.acc_stride = stride; p.tmem_cols = 4 * stride; }
    }
    p.bias = a->bias; p.row_scale = a->row_scale; p.norm_g = a->norm_g; p.scale_shift = a->scale_shift;
    p.ss_stride = a->ss_stride; p.act = a->act;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.ld_res = a->ld_res;
    p.out = a->out; p.out_f32_nchw = a->out_f32_nchw; p.ld_out = a->ld_out;
    p.OH = a->OH; p.OW = a->OW; p.oy = a->oy; p.ox = a->ox; p.sy = a->sy; p.sx = a->sx;
    p.rnorm_out = a->rnorm_out;
    p.debug = g_conv_debug;
    // streamed weights are the dominant L2->SM traffic of the wide/deep layers (every M tile re-reads the whole weight
    // matrix): pairs of CTAs in a cluster can fetch half of each chunk and multicast it.  Measured: no gain (the limit is
    // each SM's ~42 B/clk TMA ingest, which multicast does not reduce), so it is off unless DDM_CONV_DEBUG & 64 is set.
    p.cluster = (!p.b_resident && p.m_tiles >= 2 && p.n_tiles <= 2 && (p.block_n % 16) == 0 && (g_conv_debug & 64)) ? 2 : 1;
    p.pairs = (p.m_tiles + 1) / 2;
    if (pair_norm) { p.cluster = 2; p.pair_n = 1; }
    CUtensorMap tmA0, tmA1, tmW, tmOut;
    const unsigned box[5] = {64u, static_cast<unsigned>(p.bw), 1u, static_cast<unsigned>(p.bh), static_cast<unsigned>(p.bb)};
    const unsigned abox[5] = {64u, static_cast<unsigned>(p.bw), 1u, static_cast<unsigned>(p.bh + p.n_dy - 1), static_cast<unsigned>(p.bb)};
    auto encode_src = [&](CUtensorMap* tm, const void* base, int C, int ld) -> int {
        unsigned long long dims[5], str[5];
        const unsigned long long W = a->W, H = a->H, B = a->B, L = ld;
        if (a->view == 0) {           // [B,H,W,C] -> (c, x, p=1, y, b)
            dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
            str[0] = 1; str[1] = L; str[2] = L * W; str[3] = L * W; str[4] = L * W * H;
        } else {                      // [B,2H,2W,C] -> ((p2 c), x, p1, y, b); requires ld == C
            if (ld != C) return DDM_E_UNSUPPORTED;
            dims[0] = 2ull * C; dims[1] = W; dims[2] = 2; dims[3] = H; dims[4] = B;
            str[0] = 1; str[1] = 2ull * L; str[2] = 2ull * W * L; str[3] = 4ull * W * L; str[4] = 4ull * W * H * L;
        }
        return encode_bf16_map(tm, base, 5, dims, str, abox);
    };
    int r = encode_src(&tmA0, a->src0, a->C0, a->ld0);
    if (r != 0) return r;
    if (a->src1 != nullptr) {
        r = encode_src(&tmA1, a->src1, a->C1, a->ld1);
        if (r != 0) return r;
    } else {
        tmA1 = tmA0;
    }
    {
        const unsigned long long dims[2] = {static_cast<unsigned long long>(a->K_pad), static_cast<unsigned long long>(a->N_pad)};
        const unsigned long long str[2] = {1ull, static_cast<unsigned long long>(a->K_pad)};
        const unsigned wbox[2] = {64u, static_cast<unsigned>(p.block_n / ((p.cluster == 2 && !p.pair_n) ? 2 : 1))};     // multicast: half per CTA
        r = encode_bf16_map(&tmW, a->weight, 2, dims, str, wbox);
        if (r != 0) return r;
    }
    CUtensorMap tmRes;
    {
        // output (and residual, which is indexed like the output) as a 5-D map whose box is one staged 64-channel tile
        auto encode_out = [&](CUtensorMap* tm, const void* base, int ld) -> int {
            unsigned long long dims[5], str[5];
            const unsigned long long L = ld, OW = a->OW, OH = a->OH, B = a->B, W = a->W, H = a->H;
            if (!strided_out) {           // [B,OH,OW,ld] -> (c, x, 1, y, b)
                dims[0] = a->N; dims[1] = OW; dims[2] = 1; dims[3] = OH; dims[4] = B;
                str[0] = 1; str[1] = L; str[2] = L * OW; str[3] = L * OW; str[4] = L * OW * OH;
            } else {                      // [B,2H,2W,N] -> ((px c), x, py, y, b): one sub-pixel phase per launch
                dims[0] = 2ull * L; dims[1] = W; dims[2] = 2; dims[3] = H; dims[4] = B;
                str[0] = 1; str[1] = 2ull * L; str[2] = OW * L; str[3] = 2ull * OW * L; str[4] = OH * OW * L;
            }
            return encode_bf16_map(tm, base, 5, dims, str, box);
        };
        if (p.tma_store && !head) {
            r = encode_out(&tmOut, a->out, a->ld_out);
            if (r != 0) return r;
        } else {
            tmOut = tmA0;
        }
        // lean epilogue: the residual tile is TMA-loaded into the staging buffer (needs the output's sub-pixel layout
        // rule, ld_res == N for strided outputs)
        p.res_tma = 0;
        tmRes = tmOut;
        if (p.fast_epilogue && a->residual != nullptr && (!strided_out || a->ld_res == a->N)) {
            r = encode_out(&tmRes, a->residual, a->ld_res);
            if (r != 0) return r;
            p.res_tma = 1;
        }
    }
    CUtensorMap tmR1 = tmRes;
    if (shortcut) {       // the shortcut's sources as 128-pixel tile boxes (no halo); tmRes carries the first one
        auto encode_r = [&](CUtensorMap* tm, const void* base, int C, int ld) -> int {
            const unsigned long long W = a->W, H = a->H, B = a->B, L = ld;
            const unsigned long long dims[5] = {static_cast<unsigned long long>(C), W, 1ull, H, B};
            const unsigned long long str[5] = {1ull, L, L * W, L * W, L * W * H};
            return encode_bf16_map(tm, base, 5, dims, str, box);
        };
        r = encode_r(&tmRes, a->rsrc0, a->rC0, a->rld0);
        if (r != 0) return r;
        if (a->rsrc1 != nullptr) {
            r = encode_r(&tmR1, a->rsrc1, a->rC1, a->rld1);
            if (r != 0) return r;
        } else {
            tmR1 = tmRes;
        }
    }
    {
        std::lock_guard<std::mutex> lock(g_plans_mu);
        if (g_plans.size() >= kMaxPlans) g_plans.clear();
        ConvPlan& c = g_plans[a];
        std::memcpy(&c.args, a, sizeof(*a)); c.p = p; c.tmA0 = tmA0; c.tmA1 = tmA1; c.tmW = tmW; c.tmOut = tmOut; c.tmRes = tmRes; c.tmR1 = tmR1;
    }
    ddm::launch_conv(tmA0, tmA1, tmW, tmOut, tmRes, tmR1, p, g_num_sms, as_stream(stream), g_pdl);
    return finish(1);
}

int ddm_conv2d_head_supported(int N, int head_n, int H, int W) {
    if (std::getenv("DDM_NO_FUSED_HEAD") != nullptr) return 0;
    if ((N != 64 && N != 128) || head_n < 1 || head_n > 4 || H < 1 || W < 1) return 0;
    return 1;
}

int ddm_conv2d_row_norm_supported(int N) {
    if (N <= 256) return 1;                                  // one accumulator holds the row
    if (g_conv_debug & 1073741824) return 0;
    return (N <= 512 && (N % 128) == 0) ? 1 : 0;            // two 64-channel-aligned N tiles in a CTA pair
}

int ddm_conv2d_suggest_ksplit(long long rows, int N_pad, int K_pad) {
    if (!g_ready || rows < 1 || N_pad < 1 || K_pad < 64 || (g_conv_debug & 536870912)) return 1;
    const long long m_tiles = (rows + 127) / 128;
    const int n_tiles = (N_pad + 255) / 256;
    const long long tiles = m_tiles * n_tiles;
    const int stages = K_pad / 64;
    // Measured (DDIM-100, 32 x 32): +2.7 % at 16 images per GPU, +4 % at 64, -1 % at 128 when layers with up to half as many tiles
    // as SMs were split (the 8 x 8 level pays more for the extra norm launch than it gains): split only below a third.
    if (tiles * 3 > g_num_sms || stages < 8) return 1;
    int ks = static_cast<int>(g_num_sms / tiles);
    if (ks > stages / 4) ks = stages / 4;                      // at least four K chunks per range
    if (ks > 16) ks = 16;
    while (ks > 1 && (ks - 1) * ((stages + ks - 1) / ks) >= stages) --ks;     // every range gets at least one stage
    return ks < 2 ? 1 : ks;
}

int ddm_conv2d_shortcut_supported(int N, int C_in, int rC0, int rC1, int H, int W) {
    // the lean plan with a second accumulator: 64 channels (resident weights: 9 x C_in/64 + (rC0 + rC1)/64 chunks of 8 KB next to
    // >= 3 slab stages) or 128 channels (streamed weights, 2 x (128 + 128) TMEM columns)
    if (g_conv_debug & 268435456) return 0;
    if ((N != 64 && N != 128) || C_in != N || rC0 < 64 || (rC0 % 64) != 0 || (rC1 % 64) != 0 || rC0 + rC1 > 256) return 0;
    const int bw = W >= 32 ? 32 : pow2_ceil(W);             // full tiles of one image (ddm_conv2d's tile box)
    if (bw * H < 128) return 0;
    const int bh = 128 / bw;
    return ((W % bw) == 0 && (H % bh) == 0) ? 1 : 0;
}

/* debugging aid, not part of the documented ABI surface: drains the conv kernel's device-side event trace */
int ddm_debug_conv_trace(long long* host_pairs, int cap) { return ddm::conv_trace_read(host_pairs, cap); }

/* debugging aid: drains the fused linear-attention kernel's event trace as (role, tag, clock) triples */
int ddm_debug_linattn_trace(long long* host_triples, int cap) { return ddm::linattn_trace_read(host_triples, cap); }

int ddm_stem_conv(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* weight,
                  const float* bias, void* out_bf16, int B, int H, int W, int Cout, int ksize, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (in0 == nullptr || weight == nullptr || bias == nullptr || out_bf16 == nullptr || (ksize % 2) != 1) return DDM_E_BAD_ARGUMENT;
    if ((Cout % 8) != 0 || !aligned16(out_bf16)) return DDM_E_ALIGNMENT;
    if (B > 65535) return DDM_E_UNSUPPORTED;
    if (!(g_conv_debug & 134217728) && ddm::stem_umma_supported(c0 + c1 + c2, Cout, ksize, H, W)) {   // tcgen05 over an explicit im2col tile
        ddm::launch_stem_umma(in0, c0, in1, c1, in2, c2, weight, bias, out_bf16, B, H, W, Cout, g_num_sms, as_stream(stream), g_pdl);
        return finish(1);
    }
    if (ddm::stem_tc_supported(c0 + c1 + c2, Cout, ksize)) {      // tensor-core path (mma.sync implicit GEMM)
        ddm::launch_stem_tc(in0, c0, in1, c1, in2, c2, weight, bias, out_bf16, B, H, W, Cout, ksize, g_num_sms, as_stream(stream));
        return finish(1);
    }
    if (ddm::stem_smem_bytes(c0 + c1 + c2, Cout, ksize) > 200 * 1024) return DDM_E_UNSUPPORTED;
    ddm::launch_stem(in0, c0, in1, c1, in2, c2, weight, bias, out_bf16, B, H, W, Cout, ksize, as_stream(stream));
    return finish(1);
}

int ddm_head_conv1x1(const void* x_bf16, const float* weight, const float* bias, float* out_f32_nchw, int B, int HW, int C,
                     int N, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if ((C % 8) != 0 || !aligned16(x_bf16) || (C % 4) != 0) return DDM_E_ALIGNMENT;
    if (N * C * 4 > 48 * 1024) return DDM_E_UNSUPPORTED;
    const int r = ddm::launch_head_conv(x_bf16, weight, bias, out_f32_nchw, static_cast<long long>(B) * HW, C, N, HW, as_stream(stream));
    return r != 0 ? r : finish(1);
}

int ddm_sinusoidal_embedding(const float* t, float* out, int rows, int dim, float theta, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (dim < 4 || (dim % 2) != 0 || rows < 1) return DDM_E_BAD_ARGUMENT;
    ddm::launch_sinusoidal(t, out, rows, dim, theta, as_stream(stream));
    return finish(1);
}

int ddm_small_linear(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int rows, int N, int K,
                     int act_in, int act_out, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (rows < 1 || N < 1 || K < 1) return DDM_E_BAD_ARGUMENT;
    ddm::launch_small_linear(x, ldx, W, b, y, ldy, rows, N, K, act_in, act_out, as_stream(stream));
    return finish(1);
}

int ddm_row_rnorm(const void* x_bf16, int ld, float* rnorm, long long rows, int C, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if ((C % 8) != 0 || (ld % 8) != 0 || !aligned16(x_bf16)) return DDM_E_ALIGNMENT;
    ddm::launch_row_rnorm(x_bf16, ld, rnorm, rows, C, as_stream(stream));
    return finish(1);
}

int ddm_rmsnorm_act(const void* x_bf16, const float* norm_g, const float* scale_shift, long long ss_stride,
                    long long rows_per_batch, int act, const void* residual_bf16, void* out_bf16, long long rows, int C,
                    void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if ((C % 8) != 0 || !aligned16(x_bf16) || !aligned16(out_bf16) || (residual_bf16 != nullptr && !aligned16(residual_bf16)))
        return DDM_E_ALIGNMENT;
    if (rows_per_batch < 1) return DDM_E_BAD_ARGUMENT;
    ddm::launch_rmsnorm_act(x_bf16, norm_g, scale_shift, ss_stride, rows_per_batch, act, residual_bf16, out_bf16, rows, C,
                            as_stream(stream));
    return finish(1);
}

int ddm_rmsnorm_act_split(const float* partials, int ksplit, const float* bias, const float* norm_g, const float* scale_shift,
                          long long ss_stride, long long rows_per_batch, int act, const void* residual_bf16, void* out_bf16,
                          long long rows, int C, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (partials == nullptr || ksplit < 1 || rows_per_batch < 1 || rows < 1) return DDM_E_BAD_ARGUMENT;
    if ((C % 8) != 0 || !aligned16(partials) || !aligned16(out_bf16) || (residual_bf16 != nullptr && !aligned16(residual_bf16)) ||
        (bias != nullptr && !aligned16(bias)))
        return DDM_E_ALIGNMENT;
    if (C > 1024) return DDM_E_UNSUPPORTED;
    ddm::launch_rmsnorm_act_split(partials, ksplit, bias, norm_g, scale_shift, ss_stride, rows_per_batch, act, residual_bf16, out_bf16, rows,
                                  C, as_stream(stream));
    return finish(1);
}

int ddm_groupnorm_act(const void* x_bf16, const float* gamma, const float* beta, void* out_bf16, int B, int HW, int C, int groups, float eps,
                      int act, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (x_bf16 == nullptr || gamma == nullptr || beta == nullptr || out_bf16 == nullptr || B < 1 || HW < 1 || B > 0x7FFFFFF) return DDM_E_BAD_ARGUMENT;
    if (!aligned16(x_bf16) || !aligned16(out_bf16)) return DDM_E_ALIGNMENT;
    const int r = ddm::launch_groupnorm_act(x_bf16, gamma, beta, out_bf16, B, HW, C, groups, eps, act, as_stream(stream));
    return r != 0 ? DDM_E_UNSUPPORTED : finish(1);
}

int ddm_linear_attention(const void* qkv_bf16, const float* mem_kv, void* out_bf16, int B, int n, int heads, int d,
                         int n_mem, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (!aligned16(qkv_bf16) || !aligned16(out_bf16)) return DDM_E_ALIGNMENT;
    if (B > 65535 || n_mem < 0 || n_mem > 16 || (n_mem > 0 && mem_kv == nullptr)) return DDM_E_UNSUPPORTED;
    const int r = ddm::launch_linear_attention(qkv_bf16, mem_kv, nullptr, out_bf16, B, n, heads, d, n_mem, as_stream(stream));
    return r != 0 ? r : finish(1);
}

int ddm_linear_attention_bounded(const void* qkv_bf16, const float* mem_kv, const float* k_shift, void* out_bf16, int B, int n, int heads,
                                 int d, int n_mem, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (k_shift == nullptr) return DDM_E_BAD_ARGUMENT;
    if (!aligned16(qkv_bf16) || !aligned16(out_bf16)) return DDM_E_ALIGNMENT;
    if (d != 32 || B > 65535 || n_mem < 0 || n_mem > 16 || (n_mem > 0 && mem_kv == nullptr)) return DDM_E_UNSUPPORTED;
    const int r = ddm::launch_linear_attention(qkv_bf16, mem_kv, k_shift, out_bf16, B, n, heads, d, n_mem, as_stream(stream));
    return r != 0 ? r : finish(1);
}

int ddm_linear_attention_block_supported(int C, int n, int heads, int dim_head, int n_mem) {
    return ddm::linattn_fused_supported(C, n, heads, dim_head, n_mem) ? 1 : 0;
}

int ddm_linear_attention_block(const ddm_linattn_block_args* a, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (a == nullptr || a->x == nullptr || a->out == nullptr || a->w_qkv == nullptr || a->w_out == nullptr || a->bias_out == nullptr ||
        a->g_out == nullptr || a->k_shift == nullptr || a->B < 1 || (a->n_mem > 0 && a->mem_kv == nullptr) || a->x == a->out)
        return DDM_E_BAD_ARGUMENT;
    if (!ddm::linattn_fused_supported(a->C, a->n, a->heads, a->dim_head, a->n_mem)) return DDM_E_UNSUPPORTED;
    if (static_cast<long long>(a->B) * a->n > 0x7FFFFFFFll) return DDM_E_UNSUPPORTED;
    const int hid = a->heads * a->dim_head;
    CUtensorMap tmX, tmY, tmWqkv, tmWout;
    const unsigned box[2] = {64u, 128u};
    {
        const unsigned long long dims[2] = {static_cast<unsigned long long>(a->C), static_cast<unsigned long long>(a->B) * a->n};
        const unsigned long long str[2] = {1ull, static_cast<unsigned long long>(a->C)};
        int r = encode_bf16_map(&tmX, a->x, 2, dims, str, box);
        if (r != 0) return r;
        r = encode_bf16_map(&tmY, a->out, 2, dims, str, box);
        if (r != 0) return r;
    }
    {
        const unsigned long long dims[2] = {static_cast<unsigned long long>(a->C), static_cast<unsigned long long>(3 * hid)};
        const unsigned long long str[2] = {1ull, static_cast<unsigned long long>(a->C)};
        const int r = encode_bf16_map(&tmWqkv, a->w_qkv, 2, dims, str, box);
        if (r != 0) return r;
    }
    {
        const unsigned long long dims[2] = {static_cast<unsigned long long>(hid), static_cast<unsigned long long>(a->C)};
        const unsigned long long str[2] = {1ull, static_cast<unsigned long long>(hid)};
        const unsigned obox[2] = {64u, static_cast<unsigned>(a->C)};
        const int r = encode_bf16_map(&tmWout, a->w_out, 2, dims, str, obox);
        if (r != 0) return r;
    }
    ddm::launch_linattn_fused(tmX, tmY, tmWqkv, tmWout, a->bias_out, a->g_out, a->mem_kv, a->k_shift, a->B, a->n, a->C, a->n_mem,
                              g_num_sms, g_laf_trace, g_pdl, as_stream(stream));
    return finish(1);
}

int ddm_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const float* mem_k,
                  const float* mem_v, int n_mem, void* out_bf16, int B, int nq, int nk, int heads, int d, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out_bf16) || (ldq % 8) || (ldk % 8) || (ldv % 8))
        return DDM_E_ALIGNMENT;
    if (B > 65535 || heads > 65535 || n_mem < 0 || n_mem > 64) return DDM_E_UNSUPPORTED;
    if (g_tc_attention && ddm::attention_tc_supported(d, n_mem) && static_cast<long long>(B) * (nq > nk ? nq : nk) < 0x7FFFFFFFll) {
        // tcgen05 path: S = Q K^T and O = P V on the tensor cores (attention_tc.cu)
        CUtensorMap tmQ, tmK, tmV;
        const unsigned box[2] = {64u, 128u};
        const unsigned long long w = static_cast<unsigned long long>(heads) * d;
        auto enc = [&](CUtensorMap* tm, const void* base, int ld, long long rows) -> int {
            const unsigned long long dims[2] = {w, static_cast<unsigned long long>(rows)};
            const unsigned long long str[2] = {1ull, static_cast<unsigned long long>(ld)};
            return encode_bf16_map(tm, base, 2, dims, str, box);
        };
        int e = enc(&tmQ, q, ldq, static_cast<long long>(B) * nq);
        if (e == 0) e = enc(&tmK, k, ldk, static_cast<long long>(B) * nk);
        if (e == 0) e = enc(&tmV, v, ldv, static_cast<long long>(B) * nk);
        if (e != 0) return e;
        ddm::launch_attention_tc(tmQ, tmK, tmV, mem_k, mem_v, n_mem, out_bf16, B, nq, nk, heads, d, as_stream(stream));
        return finish(1);
    }
    const int r = ddm::launch_attention(q, ldq, k, ldk, v, ldv, mem_k, mem_v, n_mem, out_bf16, B, nq, nk, heads, d, as_stream(stream));
    return r != 0 ? r : finish(1);
}

int ddm_sampler_step(int kind, float* x, const float* model_out, const float* noise, long long noise_step_stride,
                     float* x_start_out, const float* coef,
                     int* step_counter, int advance, int objective, unsigned long long seed, long long numel, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if ((kind != DDM_SAMPLER_DDIM && kind != DDM_SAMPLER_DDPM) || objective < 0 || objective > 2 || numel < 1) return DDM_E_BAD_ARGUMENT;
    ddm::launch_sampler_step(kind, x, model_out, noise, noise_step_stride, x_start_out, coef, step_counter, advance, objective, seed, numel,
                             as_stream(stream));
    return finish(advance ? 2 : 1);
}

int ddm_sampler_step_learned(float* x, const float* model_out, const float* noise, long long noise_step_stride, float* x_start_out,
                             const float* coef, int* step_counter, int advance, unsigned long long seed, long long numel,
                             long long per_sample, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (x == nullptr || model_out == nullptr || coef == nullptr || step_counter == nullptr) return DDM_E_BAD_ARGUMENT;
    if (numel < 1 || per_sample < 1 || (numel % per_sample) != 0) return DDM_E_BAD_ARGUMENT;
    ddm::launch_sampler_step_learned(x, model_out, noise, noise_step_stride, x_start_out, coef, step_counter, advance, seed, numel,
                                     per_sample, as_stream(stream));
    return finish(advance ? 2 : 1);
}

int ddm_sampler_step_guided(float* x, const float* model_out, const float* noise, long long noise_step_stride, const float* guide,
                            const float* mask, const float* guide_noise, long long guide_noise_step_stride, float* x_start_out,
                            const float* coef, int* step_counter, int advance, int objective, int clip_denoised, unsigned long long seed,
                            long long numel, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    if (x == nullptr || model_out == nullptr || coef == nullptr || step_counter == nullptr || objective < 0 || objective > 2 || numel < 1)
        return DDM_E_BAD_ARGUMENT;
    if ((guide == nullptr) != (mask == nullptr)) return DDM_E_BAD_ARGUMENT;
    ddm::launch_sampler_step_guided(x, model_out, noise, noise_step_stride, guide, mask, guide_noise, guide_noise_step_stride, x_start_out,
                                    coef, step_counter, advance, objective, clip_denoised, seed, numel, as_stream(stream));
    return finish(advance ? 2 : 1);
}

int ddm_finalize(const float* x, float* y, int unnormalize, long long numel, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    ddm::launch_finalize(x, y, unnormalize, numel, as_stream(stream));
    return finish(1);
}

int ddm_select_row(const float* table, const int* step_counter, float* dst, int row_len, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    ddm::launch_select_row(table, step_counter, dst, row_len, as_stream(stream));
    return finish(1);
}

int ddm_randn(float* x, unsigned long long seed, unsigned long long stream_id, long long numel, void* stream) {
    if (!g_ready) return DDM_E_NOT_INITIALISED;
    ddm::launch_randn(x, seed, stream_id, numel, as_stream(stream));
    return finish(1);
}

}  // extern "C"
