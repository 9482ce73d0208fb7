// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), fed by TMA, with the block epilogue fused.
//
// GEMM view:  M = output pixels (tiles of 128 = bw x bh x bb box of the NHWC grid), N = C_out, K = taps x C_in.
//   A tile (128 pixels x 64 channels, bf16) : one 5-D TMA box load per (tap, 64-channel chunk); the tap offset is
//     added to the box coordinate, and out-of-bounds pixels are zero-filled by TMA, which *is* the conv padding.
//   B tile (block_n x 64, bf16)             : 2-D TMA box of the packed weight matrix [N_pad][K_pad] (K contiguous).
//   D (128 x block_n fp32)                  : TMEM accumulator, double buffered so the epilogue of tile i overlaps
//     the main loop of tile i+1.
// Warp roles: warp 0 = TMA producer (1 lane), warps 1-2 = MMA issuers (1 lane each, alternating pipeline stages;
// warp 1 also owns the TMEM allocation), warps 3..18 = epilogue:
// thread = (accumulator row, column part), so the per-pixel RMSNorm over C_out is four partial sums exchanged through
// shared memory.  Per-column epilogue vectors (bias, norm gain x (scale+1), shift) are staged in shared memory once
// per CTA; the bf16 output tile is staged in 128B-swizzled shared memory and written with TMA stores.
//
// Reference ops folded here (denoising_diffusion.py): Block.forward :113-122 (conv -> RMSNorm -> scale/shift -> SiLU),
// ResnetBlock residual add :148, Downsample :54-58 (view 1), Upsample :48-52 (4 sub-pixel phases, strided output),
// skip torch.cat :378-387 (two sources), attention pre-norm :176/:218 (row_scale), to_out + RMSNorm :169-172.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddm {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements per k-chunk = one 128-byte swizzle row
constexpr int kATileBytes = kTileM * kChunkK * 2; // 16 KiB
constexpr int kMaxTaps = 9;
constexpr int kMaxNPad = 1024;                   // widest C_out (N tiles of at most 256 columns)

struct ConvParams {
    // tile domain (pixel grid the taps are applied on) and tile box
    int B, H, W;
    int bw, bh, bb;
    int tiles_x, tiles_y, m_tiles, n_tiles, total_tiles;
    int bw_shift, bh_shift;      // log2 of bw, bh
    int tiles_pow2, tx_shift, ty_shift;   // tiles_x and tiles_y both powers of two -> shift/mask tile decode
    int block_n;                 // UMMA N: multiple of 16, <= 256
    int N;                       // real C_out
    int n_pad;                   // N rounded up to 16
    // taps grouped into "slabs": taps that share (dx, p) and have consecutive dy read one A slab of bh + n_dy - 1
    // image rows; the dy shift is a 1024B-aligned row offset into the slab (needs bb == 1 and bw % 8 == 0, otherwise
    // n_dy = 1 and every tap is its own slab).  slab_tap = position of the tap in the weight matrix's K order.
    int n_slabs, n_dy;
    int slab_dx[kMaxTaps], slab_p[kMaxTaps], slab_dy0[kMaxTaps];
    int slab_tap[kMaxTaps][3];
    int a_rows;                  // rows (pixels) of one A stage
    // dx-folded 3x3 mode (C_out = 64, tile spans the full image width, resident weights): ONE slab (dx = 0) per
    // 64-channel chunk; the three dx taps become three 64-column groups of an N = 192 MMA (weights of (dy, dx=-1|0|+1)
    // stacked along N), and the epilogue adds the groups shifted by one pixel along x (warp shuffles).  A is fetched
    // from L2 once instead of three times; fold_dyi/dxi give each tap's position (0..2) in the stacked layout.
    int fold;                    // 0 off; 3 = all three dx groups folded; 2 = [dx=-1 | dx=0] folded, dx=+1 as its own slab
    int fold_dyi[kMaxTaps], fold_dxi[kMaxTaps];
    int b_resident;              // 1: the whole weight matrix stays in shared memory for the CTA's lifetime
    int k_chunks;                // taps x 64-channel chunks
    int chunks0, chunks1;        // 64-channel chunks per tap for source 0 / source 1
    int acc_stride;              // TMEM columns between accumulator stages
    int acc_stages;              // 2, or 4 with the lean epilogue when 4 accumulators fit in TMEM: the MMA side runs
                                 // further ahead of the epilogue, hiding the accumulator hand-over latency
    int tmem_cols;               // allocated TMEM columns (power of two >= 32)
    int num_stages;              // smem ring depth
    int staging_bufs;            // output staging buffers: 1 (generic epilogue) or one per epilogue group (lean epilogue)
    int epi_groups;              // lean epilogue: 2 groups of 8 warps, or 4 groups of 4 warps (needs 4 accumulator stages)
    int fast_epilogue;           // 1: lean epilogue kernel (see conv_tc.cu), chosen by the host when its preconditions hold
    int cluster;                 // 1, or 2: CTA pairs share the (streamed) weight chunks through TMA multicast
    int pairs;                   // ceil(m_tiles / 2) when cluster == 2
    int issue_mode;              // 0: one MMA issuer thread; 1: two issuers alternating pipeline stages in token order;
                                 // 2: two issuers alternating tiles, each with its own half of the smem ring
    int debug;                   // profiling / bisection only (env DDM_CONV_DEBUG, read once in ddm_init; the bits that act inside
                                 // the kernel -- 1 2 4 128 256 4096 8192 1048576 -- need a -DDDM_CONV_DEBUG_BUILD build); bits:
                                 //   1 skip epilogue work      2 skip MMA issue        4 skip A loads
                                 //   8 generic epilogue only   32 single MMA issuer    64 2-CTA weight multicast ON
                                 //   128 device-side event trace of CTA 0 (ddm_debug_conv_trace)   256 per-tile events only
                                 //   512 no dx-folding   1024 3-group folding   2048 two accumulator stages   4096 issuer events only
                                 //   8192 no fence before the issuer token     16384 / 32768 force issuer mode 1 / 2
                                 //   262144 wide (128 x 1) tiles   1048576 no merged accumulator barrier in the folded kernels
                                 //   4194304 / 8388608 four epilogue groups everywhere / nowhere
                                 // the product path runs with 0
    int tma_store;               // 1: stage the bf16 tile in smem and TMA-store it (needs N % 64 == 0, bf16 output)
    // fused 1x1 shortcut (ResnetBlock.res_conv): res_chunks extra pipeline stages per tile (the centre-tap tiles of the shortcut's
    // sources, the first res_chunks0 of them from the first source) multiply the resident weight chunks k_chunks .. into a second
    // accumulator (TMEM columns block_n .. 2 block_n - 1), which the lean epilogue adds (+ rbias) after the activation
    int res_chunks, res_chunks0;
    const float* rbias;
    // split-K (generic kernel only): tile t = ((m_tile * ksplit) + ks) * n_tiles + n_tile; range ks takes the pipeline stages
    // [ks * ks_per, min((ks + 1) * ks_per, stages per tile)) of the tile and stores raw fp32 rows to partial + ks * partial_stride
    int ksplit, ks_per;
    float* partial;
    long long partial_stride;
    // fused head (its own instantiation of the lean kernel): the network's last 1x1 conv (final_conv, dd:343,390: C_out -> head_n <= 4
    // channels, fp32 weights [head_n][N], fp32 NCHW output) applied to the fp32 values of the tile instead of storing them
    int head_n;
    const float* head_w;
    const float* head_b;
    float* head_out;
    // pair_n (generic kernel, cluster == 2): the two N tiles of one M tile run in the two CTAs of a cluster, so that a row wider
    // than one accumulator (C_out = 512) still gets its RMSNorm in the epilogue: the CTAs exchange their per-row sums of squares
    // through distributed shared memory (remote st + remote mbarrier arrive) once per tile
    int pair_n;
    int tight_smem;              // 1: the plan only fits without the 1 KB alignment slack: the kernel requires (and checks) that
                                 // its dynamic shared memory starts 1024-byte aligned (it does when there is no static smem)
    // epilogue
    const float* bias;           // [N] or null
    const float* row_scale;      // [B*H*W] or null : v = acc * row_scale[pixel]
    const float* norm_g;         // [N] = g * sqrt(N), or null : RMSNorm over N (needs n_tiles == 1)
    const float* scale_shift;    // [Bt][2N] (scale | shift) or null
    long long ss_stride;         // elements between batch rows of scale_shift (0 = one shared row)
    int act;                     // 0 none, 1 SiLU
    const __nv_bfloat16* residual;  // added after the activation, indexed like `out`
    int ld_res;
    int res_tma;                 // 1: the lean epilogue TMA-loads the residual tile into the staging buffer (tmRes)
    void* out;
    int out_f32_nchw;            // 0: bf16 [B,OH,OW,ld_out] ; 1: fp32 [B,N,OH,OW]
    int ld_out;
    int OH, OW, oy, ox, sy, sx;  // output pixel of tile pixel (b,y,x) is (b, y*sy+oy, x*sx+ox)
    float* rnorm_out;            // [B*OH*OW] or null : 1/max(||out_row||_2, 1e-12) of the stored row
};

// tmRes: the residual tile's map, or (fused shortcut) the shortcut's first source; tmR1: its second source
void launch_conv(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW, const CUtensorMap& tmOut,
                 const CUtensorMap& tmRes, const CUtensorMap& tmR1, const ConvParams& p, int num_sms, cudaStream_t stream, bool pdl);
// shared-memory plan: returns total dynamic bytes and the stage count that fits (0 stages = does not fit)
int conv_smem_plan(const ConvParams& p, int* num_stages);
int conv_prepare_attributes();
int conv_trace_read(long long* host, int cap);   // debugging: device-side event trace (DDM_CONV_DEBUG & 128)

}  // namespace ddm
