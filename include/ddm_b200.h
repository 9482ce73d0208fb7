/*
 * ddm_b200.h -- C ABI of libddm_b200.so: the B200 (sm_100a) kernels behind the denoising hot path of
 * lbarseghyan/diffusion-models (U-Net eps-prediction forward inside the DDPM / DDIM / LDM sampling loop).
 *
 * The reference is pure Python/PyTorch and has no FFI; its "plugin boundary" for this path is the set of ATen ops
 * its modules call.  Each entry point below replaces the op sites cited next to it (paths relative to
 * /root/reference/denoising-diffusion-pytorch/denoising_diffusion/; dd = denoising_diffusion.py,
 * tc = denoising_diffusion_text_conditional.py, at = attend.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a CUDA device pointer unless stated otherwise;
 *   - activations are bf16, channels-last ([B, H, W, C], C contiguous); sampler state (x_t, eps) is fp32 NCHW
 *     exactly like the reference's tensors;
 *   - every launch goes to the caller's stream (`stream` is a cudaStream_t passed as void*), never synchronises,
 *     never allocates, and is CUDA-graph capturable;
 *   - every function returns 0 on success, a positive cudaError_t, or a negative DDM_E_* code.
 *   - there is no CPU fallback: without a B200 the library loads (so the symbols can be checked) but ddm_init fails.
 */
#ifndef DDM_B200_H_
#define DDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDM_ABI_VERSION 2

#define DDM_E_NOT_INITIALISED (-1)
#define DDM_E_BAD_ARGUMENT (-2)
#define DDM_E_UNSUPPORTED (-3)
#define DDM_E_ALIGNMENT (-4)
#define DDM_E_DRIVER (-5)
#define DDM_E_WRONG_ARCH (-6)

#define DDM_MAX_TAPS 9

int ddm_abi_version(void);
/* Must be called once per process per device before any launch: resolves cuTensorMapEncodeTiled, checks for
 * compute capability 10.x, raises the dynamic shared-memory limits. */
int ddm_init(int device);
const char* ddm_error_string(int code);
/* Kernel launches issued through this library since ddm_init (bench.py's gpu_launches claim). */
long long ddm_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * K1/K2: implicit-GEMM convolution on tcgen05 with the block epilogue fused.
 * Replaces: nn.Conv2d 3x3 in Block.proj (dd:108,114), stage-end 3x3 convs (dd:319,336), Upsample conv (dd:48-52,
 * as four sub-pixel phases), Downsample unshuffle + 1x1 (dd:54-58, view = 1), res_conv / to_qkv / to_out /
 * final_conv 1x1 (dd:134,166,169,212,213,343), nn.Linear over token matrices (tc:45-52), and the ops fused behind
 * them: RMSNorm (dd:60-67), scale/shift (dd:117-119), SiLU (dd:121), residual add (dd:148,368), torch.cat of skip
 * connections (dd:378,381,387) as a second source.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct ddm_conv_args {
    /* sources: bf16 channels-last, same pixel grid, concatenated along channels (src1 may be NULL) */
    const void* src0;
    const void* src1;
    int C0, C1;         /* channels read from each source (multiples of 8)                                    */
    int ld0, ld1;       /* elements between consecutive pixels of each source (>= C)                          */
    int view;           /* 0: source is [B,H,W,C]; 1: source is [B,2H,2W,C] read as [B,H,W,(p1),(p2 c)]:      *
                         *    tap_p selects p1, the channel axis is (p2, c) of length 2C (pixel-unshuffle)      */
    int B, H, W;        /* tile domain: one GEMM row per (b, y, x)                                             */
    int ntaps;
    int tap_dy[DDM_MAX_TAPS];
    int tap_dx[DDM_MAX_TAPS];
    int tap_p[DDM_MAX_TAPS];
    /* packed weights: bf16 [N_pad][K_pad], K contiguous, K order = (tap, source, channel) with each
     * (tap, source) segment zero-padded to a multiple of 64; N_pad = N rounded up to 16 */
    const void* weight;
    int N, N_pad, K_pad;
    /* epilogue, applied in this order */
    const float* row_scale;    /* [B*H*W] or NULL: acc *= row_scale[pixel]  (attention pre-norm, dd:176,218)   */
    const float* bias;         /* [N] or NULL                                                                   */
    const float* norm_g;       /* [N] (g * sqrt(N)) or NULL: RMSNorm over the N outputs of a pixel; N <= 256  */
    const float* scale_shift;  /* [Bt][2N] or NULL: v = v * (scale + 1) + shift                                 */
    long long ss_stride;       /* elements between batch rows of scale_shift; 0 = all rows share row 0         */
    int act;                   /* 0 = none, 1 = SiLU                                                            */
    const void* residual;      /* bf16, addressed like `out`, added last; or NULL                              */
    int ld_res;
    void* out;
    int out_f32_nchw;          /* 0: bf16 [B,OH,OW,ld_out]; 1: fp32 [B,N,OH,OW]                                 */
    int ld_out;
    int OH, OW, oy, ox, sy, sx; /* tile pixel (b,y,x) is stored at (b, y*sy+oy, x*sx+ox)                        */
    float* rnorm_out;          /* [B*OH*OW] or NULL: 1/max(||stored row||, 1e-12) for a following pre-norm     */
    /* Fused 1x1 shortcut (ResnetBlock.res_conv + the residual add, dd:134,148), or rsrc0 == NULL:
     *   out += W_r . cat(rsrc0, rsrc1)[pixel] + rbias, added last, in place of `residual` (which must be NULL).
     * rsrc*: bf16 [B,H,W,rld*], rC* channels each (multiples of 64; rsrc1 may be NULL).  The shortcut's weights
     * [N][rC0 + rC1] are appended to `weight` along K: K_pad = taps x segments x 64 + rC0 + rC1.  Runs as extra K
     * steps into a second accumulator of the same tile; see ddm_conv2d_shortcut_supported. */
    const void* rsrc0;
    const void* rsrc1;
    int rC0, rC1, rld0, rld1;
    const float* rbias;        /* [N] or NULL                                                                   */
    /* Split-K for layers with few output rows (the 4x4 / 8x8 levels at small per-GPU batches), or ksplit <= 1:
     * the K steps of every tile are divided into `ksplit` contiguous ranges, one CTA each; range i stores its raw
     * fp32 accumulator rows to partial_out[i][B*H*W][N] and nothing else happens (every epilogue field above must be
     * NULL / 0 and `out` is not written): ddm_rmsnorm_act_split sums the ranges and applies the Block epilogue. */
    int ksplit;
    float* partial_out;
    /* Fused head (the network's last 1x1 conv, final_conv dd:343,390), or head_out == NULL: instead of storing the tile,
     * head_out[b][o][y][x] = head_b[o] + sum_n head_w[o][n] * value[n] on the fp32 values the epilogue would have rounded and
     * stored (`out` is not written and may be NULL).  head_n <= 4 outputs; see ddm_conv2d_head_supported. */
    const float* head_w;       /* fp32 [head_n][N] (nn.Conv2d 1x1 layout)                                        */
    const float* head_b;       /* fp32 [head_n]                                                                  */
    float* head_out;           /* fp32 NCHW [B][head_n][H][W]                                                    */
    int head_n;
} ddm_conv_args;

int ddm_conv2d(const ddm_conv_args* args, void* stream);
/* 1 when ddm_conv2d takes the fused shortcut for a Block-epilogue conv of this shape (3x3 over C_in channels -> N with
 * RMSNorm, unit stride, H x W images), 0 otherwise (the caller then runs the shortcut as its own 1x1 conv). */
int ddm_conv2d_shortcut_supported(int N, int C_in, int rC0, int rC1, int H, int W);
/* Split factor ddm_conv2d would want for a GEMM of `rows` output rows, N_pad columns and K_pad reduction length on this
 * device: 1 when the tile count already fills the SMs, else 2..16. */
int ddm_conv2d_suggest_ksplit(long long rows, int N_pad, int K_pad);
/* 1 when ddm_conv2d can apply `norm_g` (RMSNorm over the N outputs of a pixel) in its epilogue for this N: up to 256 columns in
 * one accumulator, up to 512 (multiples of 128) as two N tiles in a CTA pair exchanging sums of squares over distributed
 * shared memory (not together with rnorm_out).  Otherwise: ddm_conv2d without norm_g, then ddm_rmsnorm_act. */
int ddm_conv2d_row_norm_supported(int N);
/* 1 when ddm_conv2d can fuse a head of head_n outputs behind a Block-epilogue conv with N channels on H x W images. */
int ddm_conv2d_head_supported(int N, int head_n, int H, int W);
/* Debugging aid: with DDM_CONV_DEBUG & 128 the conv kernel records (tag, clock64) pairs from CTA 0; this drains them to
 * host memory (synchronises the device) and returns the number of pairs.  Not used by the product path. */
int ddm_debug_conv_trace(long long* host_pairs, int cap);

/* ---------------------------------------------------------------------------------------------------------------
 * K3: stem convolution (dd:262,356; ic:46-54).  Direct k x k conv (k = 7, pad 3) over the channel-concatenation
 * of up to three fp32 NCHW tensors, written as bf16 channels-last.  weight: fp32 [k*k*Cin][Cout] (tap-major).
 * ------------------------------------------------------------------------------------------------------------- */
int ddm_stem_conv(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* weight,
                  const float* bias, void* out_bf16, int B, int H, int W, int Cout, int ksize, void* stream);

/* Head convolution (final_conv 1x1, dd:343,390): bf16 channels-last [B*HW, C] -> fp32 NCHW [B, N, HW], N in
 * {1,2,3,4,6,8}; weight fp32 [N][C] (nn.Conv2d layout), bias fp32 [N].  HBM-bound; no weight rounding. */
int ddm_head_conv1x1(const void* x_bf16, const float* weight, const float* bias, float* out_f32_nchw, int B, int HW, int C,
                     int N, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K5: time-conditioning path (dd:77-84 sinusoid, dd:280-285 time_mlp, dd:127-130 per-block SiLU -> Linear,
 * tc:146-152 text concat).  Built from two tiny kernels.
 * ------------------------------------------------------------------------------------------------------------- */
/* out[r, :] = cat(sin(t_r f), cos(t_r f)), f_j = exp(-j ln(theta)/(half-1)); t is fp32 [rows] */
int ddm_sinusoidal_embedding(const float* t, float* out, int rows, int dim, float theta, void* stream);
/* y[r, n] = act_out( sum_k act_in(x[r, k]) W[n, k] + b[n] ); W fp32 [N][K] (nn.Linear layout).
 * act codes: 0 none, 1 SiLU, 2 exact-erf GELU.  x has row stride ldx, y has row stride ldy. */
int ddm_small_linear(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int rows, int N, int K,
                     int act_in, int act_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K4: row norms (dd:60-67).  rnorm[m] = 1/max(||x[m,:]||_2, 1e-12) over C bf16 channels (row stride ld).
 * ------------------------------------------------------------------------------------------------------------- */
int ddm_row_rnorm(const void* x_bf16, int ld, float* rnorm, long long rows, int C, void* stream);
/* Unfused Block tail for C_out > 256: out = [ +residual ] act( RMSNorm(x) * (scale+1) + shift ), all bf16 rows of
 * length C (dd:115-122,148; tc:27-36 RMSNorm1D).  rows_per_batch maps a row to its scale_shift row. */
int ddm_rmsnorm_act(const void* x_bf16, const float* norm_g, const float* scale_shift, long long ss_stride,
                    long long rows_per_batch, int act, const void* residual_bf16, void* out_bf16, long long rows, int C,
                    void* stream);
/* The same tail over split-K partial sums (ddm_conv_args.ksplit): x = bias + sum_i partials[i][row][:], fp32 [ksplit][rows][C];
 * norm_g may be NULL (plain conv + bias). */
int ddm_rmsnorm_act_split(const float* partials, int ksplit, const float* bias, const float* norm_g, const float* scale_shift,
                          long long ss_stride, long long rows_per_batch, int act, const void* residual_bf16, void* out_bf16,
                          long long rows, int C, void* stream);

/* GroupNorm (+ swish) of the VAE decoder (latent-diffusion/ldm/modules/diffusionmodules/model.py:55-56 Normalize =
 * GroupNorm(32, C, eps=1e-6), and the x*sigmoid(x) behind it at :118-119,127-128,573-574): x, out bf16 [B, HW, C]
 * channels-last, `groups` groups of C/groups consecutive channels, statistics over one image; act 0 = none, 1 = swish. */
int ddm_groupnorm_act(const void* x_bf16, const float* gamma, const float* beta, void* out_bf16, int B, int HW, int C, int groups, float eps,
                      int act, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K6: linear attention core (dd:178-192).  qkv: bf16 [B, n, 3*heads*d] (q | k | v, head-major channels),
 * mem_kv: fp32 [2][heads][d][n_mem] (dd:163); out: bf16 [B, n, heads*d].  d must be 16, 32 or 64.
 * ------------------------------------------------------------------------------------------------------------- */
int ddm_linear_attention(const void* qkv_bf16, const float* mem_kv, void* out_bf16, int B, int n, int heads, int d,
                         int n_mem, void* stream);
/* The same with a caller-supplied shift for the softmax over the tokens (d = 32 only): k_shift fp32 [heads*d] must bound
 * k[token][channel] from above for every token (and be >= the channel's memory keys) -- e.g. ||w_c||_2 when the tokens entering
 * to_qkv are unit vectors (the block's RMSNorm).  Saves the max pass over k; the result is the same softmax. */
int ddm_linear_attention_bounded(const void* qkv_bf16, const float* mem_kv, const float* k_shift, void* out_bf16, int B, int n,
                                 int heads, int d, int n_mem, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K6f: the whole LinearAttention block as ONE tcgen05 / TMA kernel (dd:173-193 plus the caller's residual dd:368,383):
 *     out = RMSNorm( W_out . LinAttn( W_qkv . RMSNorm(x) ) + bias_out ) * g_out + x
 * x, out: bf16 [B, n, C] (out must not alias x).  w_qkv: bf16 [3*heads*dim_head][C] (rows q | k | v, head-major) with
 * the pre-norm gain g*sqrt(C) folded into its columns; w_out: bf16 [C][heads*dim_head]; g_out = g*sqrt(C) of the
 * output norm; mem_kv: fp32 [2][heads][dim_head][n_mem]; k_shift: fp32 [heads*dim_head], an upper bound of
 * k[c][token] per channel (>= ||w_k row c||_2 and >= the channel's memory keys): the softmax over the tokens (dd:185)
 * is evaluated as exp(k - k_shift[c]) / sum, which is the same value for any shift and cannot overflow for a bound.
 * Replaces four launches (row norm, to_qkv, ddm_linear_attention, to_out + RMSNorm + residual) and their HBM round
 * trips.  ddm_linear_attention_block_supported() says whether a shape is covered (C = 64, heads = 4, dim_head = 32,
 * n a multiple of 128); other shapes use the unfused entry points above.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct ddm_linattn_block_args {
    const void* x;
    void* out;
    int B, n, C;
    const void* w_qkv;
    const void* w_out;
    const float* bias_out;
    const float* g_out;
    const float* mem_kv;
    const float* k_shift;
    int heads, dim_head, n_mem;
} ddm_linattn_block_args;
int ddm_linear_attention_block_supported(int C, int n, int heads, int dim_head, int n_mem);
int ddm_linear_attention_block(const ddm_linattn_block_args* args, void* stream);
/* Debugging aid (env DDM_LAF_TRACE=1): drains CTA 0's (role, tag, clock64) event triples; synchronises the device. */
int ddm_debug_linattn_trace(long long* host_triples, int cap);

/* ---------------------------------------------------------------------------------------------------------------
 * K7/K8: softmax attention (dd:220-228 + at:109-124; tc:66-77).  q: bf16 rows [B*nq] with row stride ldq, head h at
 * column h*d; k, v likewise over [B*nk] rows; optional learned memory rows mem_k/mem_v fp32 [heads][n_mem][d] are
 * prepended to the keys/values (dd:223-224).  out: bf16 [B*nq][heads*d].  scale = d^-0.5.
 * d = 32 / 64 / 128 run on tcgen05 (S = Q K^T and O = P V in TMEM, attention_tc.cu); this also serves the single-head
 * AttnBlock of the VAE decoder (ldm/modules/diffusionmodules/model.py:190-215, d = C).  d = 16 uses CUDA cores.
 * ------------------------------------------------------------------------------------------------------------- */
int ddm_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const float* mem_k,
                  const float* mem_v, int n_mem, void* out_bf16, int B, int nq, int nk, int heads, int d, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K9: per-timestep sampler update, one fused elementwise kernel (dd:603-626 model_predictions, dd:684-701 DDIM,
 * dd:628-645 + dd:594-601 ancestral DDPM).  State tensors are fp32 NCHW with `numel` elements.  The per-step
 * coefficients live in a device table so that one captured CUDA graph can be replayed for every step:
 * row s of `coef` (8 floats) is selected by step_counter[0] (device int32[2] = {step, call epoch}), which the kernel
 * increments when `advance` != 0; step_counter[1] only salts the in-kernel noise stream so that replaying the same graph
 * for a new sampling call draws new noise.
 *   DDIM row : { sqrt_recip_acp[t], sqrt_recipm1_acp[t], sqrt(acp[t_next]), c, sigma, is_last, sqrt_acp[t], sqrt_1m_acp[t] }
 *   DDPM row : { sqrt_recip_acp[t], sqrt_recipm1_acp[t], coef1[t], coef2[t], exp(0.5 logvar[t]) or 0 at t=0, 0,
 *                sqrt_acp[t], sqrt_1m_acp[t] }
 * objective: 0 pred_noise, 1 pred_x0, 2 pred_v.  noise: fp32, the draw of step s starts at noise + s*noise_step_stride
 * (injected noise, parity mode), or NULL to draw N(0,1) in-kernel from Philox4x32-10 keyed by (seed, step, element);
 * x_start_out may be NULL.
 * ------------------------------------------------------------------------------------------------------------- */
#define DDM_SAMPLER_DDIM 0
#define DDM_SAMPLER_DDPM 1
int ddm_sampler_step(int kind, float* x, const float* model_out, const float* noise, long long noise_step_stride,
                     float* x_start_out, const float* coef, int* step_counter, int advance, int objective, unsigned long long seed,
                     long long numel, void* stream);
/* Ancestral step with a learned variance (learned_gaussian_diffusion.py:91-111 p_mean_variance + dd:638-645 p_sample):
 * model_out is [B, 2C, H, W] fp32 = (pred_noise | v), v in [-1, 1] interpolates the log-variance between
 * posterior_log_variance_clipped[t] and log(betas[t]).  per_sample = C*H*W, numel = B*per_sample.
 *   row : { sqrt_recip_acp[t], sqrt_recipm1_acp[t], coef1[t], coef2[t], t > 0 ? 1 : 0, min_log[t], max_log[t], 0 }
 * noise / step_counter / advance / seed as in ddm_sampler_step. */
int ddm_sampler_step_learned(float* x, const float* model_out, const float* noise, long long noise_step_stride, float* x_start_out,
                             const float* coef, int* step_counter, int advance, unsigned long long seed, long long numel,
                             long long per_sample, void* stream);
/* Guided DDIM step (dd:710-777 ddim_sample_guided): x0 / eps as model_predictions returns them with
 * clip_x_start = clip_denoised and WITHOUT re-deriving eps, the DDIM update of the DDIM row above, then (not on the
 * last step) the known region is overwritten with the guide noised to step t (dd:747-749):
 *   x = x * mask + (sqrt_acp[t] * guide + sqrt_1m_acp[t] * z_g) * (1 - mask)
 * guide, mask: fp32, `numel` elements each (mask already broadcast), or both NULL (plain unguided loop);
 * guide_noise: injected z_g per step (parity mode) or NULL to draw it in-kernel. */
int ddm_sampler_step_guided(float* x, const float* model_out, const float* noise, long long noise_step_stride, const float* guide,
                            const float* mask, const float* guide_noise, long long guide_noise_step_stride, float* x_start_out,
                            const float* coef, int* step_counter, int advance, int objective, int clip_denoised, unsigned long long seed,
                            long long numel, void* stream);
/* y = (x + 1) * 0.5 (utils.py:48-49) or a plain copy when unnormalize == 0 */
int ddm_finalize(const float* x, float* y, int unnormalize, long long numel, void* stream);
/* dst[i] = src[(*step_counter) * row_len + i]: selects the current step's row of a precomputed per-step table */
int ddm_select_row(const float* table, const int* step_counter, float* dst, int row_len, void* stream);
/* x ~ N(0,1) from the library's Philox stream (throughput mode x_T, dd:651,676) */
int ddm_randn(float* x, unsigned long long seed, unsigned long long stream_id, long long numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDM_B200_H_ */
