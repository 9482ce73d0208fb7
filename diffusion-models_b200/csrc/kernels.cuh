// Internal launcher declarations shared between the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>

namespace ddm {

constexpr int DDM_KIND_DDIM = 0;
constexpr int DDM_KIND_DDPM = 1;

// small_kernels.cu
int stem_smem_bytes(int Cin, int Cout, int ks);
int stem_prepare_attributes();
void launch_stem(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                 void* out, int B, int H, int W, int Cout, int ks, cudaStream_t s);
void launch_sinusoidal(const float* t, float* out, int rows, int dim, float theta, cudaStream_t s);
void launch_small_linear(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int rows, int N, int K,
                         int act_in, int act_out, cudaStream_t s);
void launch_row_rnorm(const void* x, int ld, float* rn, long long rows, int C, cudaStream_t s);
void launch_rmsnorm_act(const void* x, const float* g, const float* ss, long long ss_stride, long long rows_per_batch, int act,
                        const void* res, void* out, long long rows, int C, cudaStream_t s);
void launch_rmsnorm_act_split(const float* partials, int ksplit, const float* bias, const float* g, const float* ss, long long ss_stride,
                              long long rows_per_batch, int act, const void* res, void* out, long long rows, int C, cudaStream_t s);
int launch_groupnorm_act(const void* x, const float* gamma, const float* beta, void* out, int B, int HW, int C, int G, float eps, int act,
                         cudaStream_t s);
int launch_head_conv(const void* x, const float* w, const float* bias, float* out, long long rows, int C, int N, int HW,
                     cudaStream_t s);
void launch_sampler_step(int kind, float* x, const float* mo, const float* noise, long long noise_stride, float* x0_out, const float* coef, int* step_counter,
                         int advance, int objective, unsigned long long seed, long long numel, cudaStream_t s);
void launch_sampler_step_learned(float* x, const float* mo, const float* noise, long long noise_stride, float* x0_out, const float* coef,
                                 int* step_counter, int advance, unsigned long long seed, long long numel, long long per_sample, cudaStream_t s);
void launch_sampler_step_guided(float* x, const float* mo, const float* noise, long long noise_stride, const float* guide, const float* mask,
                                const float* gnoise, long long gnoise_stride, float* x0_out, const float* coef, int* step_counter, int advance,
                                int objective, int clip, unsigned long long seed, long long numel, cudaStream_t s);
void launch_finalize(const float* x, float* y, int unnorm, long long numel, cudaStream_t s);
void launch_select_row(const float* table, const int* step_counter, float* dst, int row_len, cudaStream_t s);
void launch_randn(float* x, unsigned long long seed, unsigned long long sid, long long numel, cudaStream_t s);

// stem_umma.cu
bool stem_umma_supported(int Cin, int Cout, int ks, int H, int W);
int stem_umma_prepare_attributes();
void launch_stem_umma(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                      void* out, int B, int H, int W, int Cout, int num_sms, cudaStream_t s, bool pdl);
// stem_tc.cu
bool stem_tc_supported(int Cin, int Cout, int ks);
int stem_tc_prepare_attributes();
void launch_stem_tc(const float* in0, int c0, const float* in1, int c1, const float* in2, int c2, const float* w, const float* b,
                    void* out, int B, int H, int W, int Cout, int ks, int num_sms, cudaStream_t s);

// linattn_tc.cu
void launch_linattn32_tc(const void* qkv, const float* mem_kv, const float* k_shift, void* out, int B, int n, int heads, int n_mem,
                         cudaStream_t s);

// linattn_fused.cu (CUtensorMap-taking launcher declared in api.cu, which includes <cuda.h>)
int linattn_fused_prepare_attributes();
bool linattn_fused_supported(int C, int n, int heads, int d, int n_mem);

// attention.cu
int attention_prepare_attributes();
int launch_linear_attention(const void* qkv, const float* mem_kv, const float* k_shift, void* out, int B, int n, int heads, int d, int n_mem,
                            cudaStream_t s);
int launch_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const float* mem_k, const float* mem_v,
                     int n_mem, void* out, int B, int nq, int nk, int heads, int d, cudaStream_t s);

}  // namespace ddm
