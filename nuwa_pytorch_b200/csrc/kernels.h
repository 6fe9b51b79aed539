// Internal (C++) declarations shared between the .cu translation units and the C-ABI shim (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace nuwa {

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_GLU = 2, ACT_GEGLU = 3 };

struct GemmParams {
  int M, N, K, ldw;
  int num_m_tiles, num_n_tiles, k_blocks;
  // epilogue
  const float* bias;
  const float* residual;
  float* out_f32;
  bf16* out_bf16;
  int ld_out, ld_res, act;
  // implicit-GEMM convolution geometry (output pixels)
  int conv;
  int B, H, W;
  int tw, th, tb;
  int tiles_x, tiles_y;
  int cin_blocks;
  int8_t tap_map[16];
  int8_t tap_dx[16];
  int8_t tap_dy[16];
  // split-K (weight-gradient GEMMs: small output, long contraction): item = tile * splits + split
  int splits, kb_per_split, atomic_out;
  int tma_out;  // 0: per-thread stores, 1: TMA tile stores, 2: TMA fp32 reduce-add (split-K)
  // 1: both operands are stored with the CONTRACTION as the slow dimension (A as [K][M], W as [K][N], rows = contraction
  // index): weight-gradient products dW = dY^T X read dY and X exactly as the forward pass left them (no transposes)
  int mn_major;
};

typedef nuwa_attn_params AttnParams;
typedef nuwa_ln_params LnParams;
typedef nuwa_embed_params EmbedParams;

int device_sm_count();

// gemm_tcgen05.cu
int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
              const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act, int force_bn,
              cudaStream_t stream, int splits = 1);
int gemm_bf16_tn_splitk(const void* At, int lda, const void* Wt, int ldw, int M, int N, int K, float* out_f32, int ld_out,
                        int splits, int force_bn, cudaStream_t stream);
int gemm_skinny(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act,
                cudaStream_t stream);  // gemm_skinny.cu
int conv2d_nhwc_bf16(const void* x, const void* w, int B, int Hin, int Win, int Cin, int Cout, int ksize, int stride,
                     const float* bias, const float* residual, float* out_f32, void* out_bf16, int act, int force_bn,
                     cudaStream_t stream);


void* gemm_prof_open();
void gemm_prof_attach(void* h);
int gemm_prof_collect(void* h, double* flops, float* ms);
double gemm_prof_bytes(void* h);
void gemm_prof_close(void* h);

// attention.cu
int attn_sparse3dna(const AttnParams& p, cudaStream_t s);
int attn_dense(const AttnParams& p, cudaStream_t s);
int attn_3dna_halo(const AttnParams& p, cudaStream_t stream);           // attention_3dna_halo.cu
int attn_3dna_umma(const AttnParams& p, cudaStream_t stream);           // attention_3dna_umma.cu
int attn_cross2dna_umma(const AttnParams& p, cudaStream_t stream);      // attention_3dna_umma.cu
int attn_dense_q1(const AttnParams& p, int nk, cudaStream_t stream);   // attention_q1.cu
int attn_dense_q1_bwd(const AttnParams& p, int nk, const void* dO, long long do_bs, void* dq, long long dq_bs, float* dk, float* dv,
                      long long dkv_bs, int dkv_rs, float* dnull_k, float* dnull_v, cudaStream_t stream);
int attn_dense_bwd_fused(const AttnParams& p, int nk, const void* dO, long long do_bs, int do_rs, void* Pp, void* dS, int jp,
                         float* dtalk, float out_scale, cudaStream_t stream);   // attention_dense_bwd.cu
int attn_3dna_umma_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t stream);
int attn_cross2dna_umma_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t stream);
int attn_cross2dna_umma_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                               cudaStream_t stream);                    // attention_3dna_umma.cu
int attn_3dna_umma_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                          cudaStream_t stream);                         // attention_3dna_umma.cu
int attn_dense_mma(const AttnParams& p, int nk, void* vT_ws, cudaStream_t stream);  // attention_mma.cu
int attn_dense_x64(const AttnParams& p, int nk, cudaStream_t stream);                // attention_x64.cu
int attn_dense_pres(const AttnParams& p, int nk, cudaStream_t stream);               // attention_dense_pres.cu
int attn_cross2dna(const AttnParams& p, cudaStream_t s);
// norm.cu
int sandwich_ln(const LnParams& p, cudaStream_t stream);
int stable_ln(const float* a, const float* b2, const float* w, const float* bias, float* out_f32, void* out_bf16,
              int rows, int D, cudaStream_t stream);
// token_ops.cu
int embed_tokens(const EmbedParams& p, cudaStream_t stream);
int rotary_to_bf16(const float* qkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                   cudaStream_t stream);
int cross_entropy_mean(const float* logits, int ld, const long long* target, float* row_loss, float* out, int rows,
                       int V, cudaStream_t stream);
int sample_topk_gumbel(const float* cond, const float* uncond, const float* noise, long long* out, float* guided_out,
                       int B, int V, int k, float cond_scale, float temperature, cudaStream_t stream);
int sample_topk_gumbel_at(const float* cond, const float* uncond, const float* noise, long long* out, long long out_bs,
                          const int* step_ptr, int B, int V, int k, float cond_scale, float temperature,
                          cudaStream_t stream);
int cache_append(const void* row, void* cache, long long cache_bs, int width, int B, const int* t_ptr, cudaStream_t stream);
int step_increment(int* t_ptr, cudaStream_t stream);
// decode_stack.cu
int decode_stack(const nuwa_decode_params& p, int cooperative, cudaStream_t stream);
// optim.cu
int sqnorm_f32(const float* x, long long n, float* partials, int nparts, float* out, int accumulate, cudaStream_t stream);
int adamw_step(const nuwa_adamw_params& a, cudaStream_t stream);
int recon_loss_f32(const float* a, const float* b, long long n, int l2, float* partials, int nparts, float* out,
                   cudaStream_t stream);
// vae_ops.cu
int nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, cudaStream_t stream);
int nhwc_to_nchw_f32(const void* in, int in_is_bf16, float* out, int B, int C, int H, int W, cudaStream_t stream);
int im2col_nchw_f32(const float* img, void* out, int B, int C, int H, int W, int KS, int Kpad, cudaStream_t stream);
int groupnorm_nhwc(const float* x, const float* w, const float* bias, float* stats_ws, void* out_bf16, float* out_f32,
                   int B, int HW, int C, int G, int leaky, cudaStream_t stream);
int upsample2x_nhwc_bf16(const void* in, void* out, int B, int H, int W, int C, cudaStream_t stream);
int vae_attn_prep(const float* qkv, void* out, int B, int n, int inner, cudaStream_t stream);
int vq_argmax(const float* x, const float* code, const float* code_sq, long long* out, int M, int Kc, int D, int cosine,
              cudaStream_t stream);
size_t vq_argmax_tc_workspace(int M, int Kc, int D);
int vq_argmax_tc(const float* x, const float* code, const float* code_sq, const void* code_bf16, const float* emax,
                 long long* out, int M, int Kc, int D, int cosine, void* workspace, size_t ws_bytes, cudaStream_t stream);
int split3_f32_bf16(const float* x, long long ld, void* out, long long rows, int K, cudaStream_t stream);
size_t linear_f32x3_workspace(int M, int K);
int linear_f32x3(const float* x, long long ldx, const void* w3, int M, int N, int K, const float* bias, float* out,
                 void* out_bf16, int ld_out, void* workspace, size_t ws_bytes, cudaStream_t stream);
int gather_rows(const float* table, const long long* idx, void* out_bf16, float* out_f32, long long M, int D,
                cudaStream_t stream);
int conv1x1_nhwc_to_nchw(const void* x, const float* w, const float* bias, float* out, int B, int HW, int C, int Cout,
                         cudaStream_t stream);


// bgemm.cu
int bgemm(const nuwa_bgemm_params& p, cudaStream_t stream);
// backward.cu
int ln_bwd_grid(int rows);
int ln_bwd(const nuwa_lnbwd_params& p, cudaStream_t stream);
int reduce_partials(const float* part, int nparts, int D, float* o0, float* o1, float* o2, cudaStream_t stream);
int transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int R, int C, cudaStream_t stream);
int geglu_fwd(const void* h, void* g, long long M, int ip, cudaStream_t stream);
int geglu_bwd(const void* dg, const void* h, void* dh, long long M, int ip, cudaStream_t stream);
int ce_bwd(const float* logits, int ld, const long long* target, const float* gscale, void* dlogits, int ld_out, int rows,
           int V, cudaStream_t stream);
int embed_bwd(const nuwa_embed_bwd_params& p, cudaStream_t stream);
int rotary_bwd_to_bf16(const float* dqkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                       cudaStream_t stream);
int add_rows_f32(float* dst, long long ld_dst, const float* src, long long ld_src, const int* map, int rows, int cols,
                 int accumulate, cudaStream_t stream);
int attn_bwd_rows(const nuwa_attn_rows_params& p, cudaStream_t stream);
int kv_full_build(const void* k, const void* v, long long kv_bs, int kv_rs, const float* null_k, const float* null_v,
                  void* kfull, void* vfull, int B, int nk, int jp, int inner, cudaStream_t stream);
int kv_full_split(const float* dkfull, const float* dvfull, float* dnull_k, float* dnull_v, void* dk16, void* dv16,
                  float* dk32, float* dv32, long long o_bs, int o_rs, int B, int nk, int jp, int inner,
                  cudaStream_t stream);
int mask_scores(float* S, const unsigned char* mask, int mask_bs, int B, int H, int nq, int jp, int nk, int has_null,
                cudaStream_t stream);
// attention.cu (gather-attention backward)
int attn3dna_bwd_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                        cudaStream_t s);
int attn3dna_bwd_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t s);
int attn3dna_bwd_dkdv(const AttnParams& p, const void* dO, long long do_bs, int do_rs, const void* dS, const void* Pp, int jp,
                      void* dk, void* dv, long long dkv_bs, int dkv_rs, cudaStream_t stream);
int attnx2_bwd_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                      cudaStream_t s);
int attnx2_bwd_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t s);
int attnx2_bwd_dkdv(const AttnParams& p, int nk, const void* dO, long long do_bs, int do_rs, const void* dS, const void* Pp,
                    int jp, const float* base_k, const float* base_v, long long base_bs, int base_rs, void* dk, void* dv,
                    long long dkv_bs, int dkv_rs, cudaStream_t stream);
int attn_bwd_first_key(const void* q, long long q_bs, int q_rs, const void* dO, long long do_bs, int do_rs, const void* dS,
                       const void* Pp, int jp, int B, int H, int dh, int nq, float* out_k, float* out_v, long long ok_bs,
                       cudaStream_t stream);
int attn3dna_bwd_first_key_finalize(const float* tmp_k, const float* tmp_v, const void* dO_bos, long long do_bs, void* dqkv,
                                    long long dqkv_bs, int inner, int B, cudaStream_t stream);

}  // namespace nuwa
