// C-ABI shim: extern "C" entry points declared in include/nuwa_b200.h.  Plain pointers and sizes
// only -- no torch types cross this boundary.
#include "../../include/nuwa_b200.h"
#include "kernels.h"

namespace nuwa {
unsigned long long g_launch_count = 0;
}

using namespace nuwa;
#define S(x) reinterpret_cast<cudaStream_t>(x)

extern "C" {

const char* nuwa_strerror(int code) {
  switch (code) {
    case NUWA_OK: return "ok";
    case NUWA_ERR_INVALID: return "invalid argument or unsupported shape";
    case NUWA_ERR_CUDA: return "CUDA runtime error (kernel launch failed)";
    case NUWA_ERR_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    case NUWA_ERR_WORKSPACE: return "workspace too small";
  }
  return "unknown error";
}
int nuwa_abi_version(void) { return 1; }
unsigned long long nuwa_launch_count(void) { return g_launch_count; }

int nuwa_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act,
                   int force_bn, void* stream) {
  return gemm_bf16(A, lda, W, ldw, M, N, K, bias, residual, ld_res, out_f32, out_bf16, ld_out, act, force_bn,
                   S(stream));
}
int nuwa_conv2d_nhwc_bf16(const void* x, const void* w, int B, int Hin, int Win, int Cin, int Cout, int ksize,
                          int stride, const float* bias, const float* residual, float* out_f32, void* out_bf16,
                          int act, int force_bn, void* stream) {
  return conv2d_nhwc_bf16(x, w, B, Hin, Win, Cin, Cout, ksize, stride, bias, residual, out_f32, out_bf16, act,
                          force_bn, S(stream));
}

int nuwa_sandwich_ln(const nuwa_ln_params* p, void* stream) { return p ? sandwich_ln(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_stable_ln(const float* a, const float* b2, const float* w, const float* bias, float* out_f32, void* out_bf16,
                   int rows, int D, void* stream) {
  return stable_ln(a, b2, w, bias, out_f32, out_bf16, rows, D, S(stream));
}
int nuwa_attn_sparse3dna(const nuwa_attn_params* p, void* vt_workspace, void* stream) {
  (void)vt_workspace;  // kept in the signature for ABI stability (was the workspace of a removed kernel variant)
  if (!p) return NUWA_ERR_INVALID;
  return attn_sparse3dna(*p, S(stream));
}
int nuwa_attn_sparse3dna_halo(const nuwa_attn_params* p, void* stream) {
  return p ? attn_3dna_halo(*p, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_sparse3dna_umma(const nuwa_attn_params* p, void* stream) {
  return p ? attn_3dna_umma(*p, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_cross2dna_umma(const nuwa_attn_params* p, void* stream) {
  return p ? attn_cross2dna_umma(*p, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_dense(const nuwa_attn_params* p, void* vt_workspace, void* stream) {
  if (!p) return NUWA_ERR_INVALID;
  if (p->nq >= 16) {  // 8 x 64 heads, learned null key / mask / talking heads: two-pass 64-query tensor-core kernel
    const int rc = attn_dense_x64(*p, p->jmax - (p->null_k != nullptr ? 1 : 0), S(stream));
    if (rc != NUWA_ERR_INVALID) return rc;
  }
  if (vt_workspace != nullptr && p->nq >= 8) {
    const int nk = p->jmax - (p->null_k != nullptr ? 1 : 0);
    const int rc = attn_dense_mma(*p, nk, vt_workspace, S(stream));
    if (rc != NUWA_ERR_INVALID) return rc;  // unsupported shape -> generic kernel below
  }
  return attn_dense(*p, S(stream));
}
int nuwa_attn_dense_pres(const nuwa_attn_params* p, void* stream) {
  if (!p) return NUWA_ERR_INVALID;
  return attn_dense_pres(*p, p->jmax - (p->null_k != nullptr ? 1 : 0), S(stream));
}
int nuwa_attn_cross2dna(const nuwa_attn_params* p, void* stream) { return p ? attn_cross2dna(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_embed_tokens(const nuwa_embed_params* p, void* stream) { return p ? embed_tokens(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_rotary_to_bf16(const float* qkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                        void* stream) {
  return rotary_to_bf16(qkv, out, inv_freq, rows, n, H, dh, rot, S(stream));
}
int nuwa_cross_entropy_mean(const float* logits, int ld, const long long* target, float* row_loss, float* out, int rows,
                            int V, void* stream) {
  return cross_entropy_mean(logits, ld, target, row_loss, out, rows, V, S(stream));
}
int nuwa_sample_topk_gumbel(const float* cond, const float* uncond, const float* noise, long long* out,
                            float* guided_out, int B, int V, int k, float cond_scale, float temperature, void* stream) {
  return sample_topk_gumbel(cond, uncond, noise, out, guided_out, B, V, k, cond_scale, temperature, S(stream));
}
int nuwa_sample_topk_gumbel_at(const float* cond, const float* uncond, const float* noise, long long* out,
                               long long out_bs, const int* step_ptr, int B, int V, int k, float cond_scale,
                               float temperature, void* stream) {
  return sample_topk_gumbel_at(cond, uncond, noise, out, out_bs, step_ptr, B, V, k, cond_scale, temperature, S(stream));
}
int nuwa_cache_append(const void* row, void* cache, long long cache_bs, int width, int B, const int* t_ptr, void* stream) {
  return cache_append(row, cache, cache_bs, width, B, t_ptr, S(stream));
}
int nuwa_step_increment(int* t_ptr, void* stream) { return step_increment(t_ptr, S(stream)); }
int nuwa_decode_stack(const nuwa_decode_params* p, int cooperative, void* stream) {
  return p ? decode_stack(*p, cooperative, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_sqnorm_f32(const float* x, long long n, float* partials, int nparts, float* out, int accumulate, void* stream) {
  return sqnorm_f32(x, n, partials, nparts, out, accumulate, S(stream));
}
int nuwa_recon_loss_f32(const float* a, const float* b, long long n, int l2, float* partials, int nparts, float* out,
                        void* stream) {
  return recon_loss_f32(a, b, n, l2, partials, nparts, out, S(stream));
}
int nuwa_adamw_step(const nuwa_adamw_params* a, void* stream) { return a ? adamw_step(*a, S(stream)) : NUWA_ERR_INVALID; }
void nuwa_struct_sizes_optim(int* out2) {
  out2[0] = (int)sizeof(nuwa_opt_chunk);
  out2[1] = (int)sizeof(nuwa_adamw_params);
}
void nuwa_struct_sizes_decode(int* out2) {
  out2[0] = (int)sizeof(nuwa_decode_sub);
  out2[1] = (int)sizeof(nuwa_decode_params);
}
int nuwa_split3_f32_bf16(const float* x, long long ld, void* out, long long rows, int K, void* stream) {
  return split3_f32_bf16(x, ld, out, rows, K, S(stream));
}
unsigned long long nuwa_linear_f32x3_workspace(int M, int K) { return linear_f32x3_workspace(M, K); }
int nuwa_linear_f32x3(const float* x, long long ldx, const void* w3, int M, int N, int K, const float* bias, float* out,
                      void* out_bf16, int ld_out, void* workspace, unsigned long long workspace_bytes, void* stream) {
  return linear_f32x3(x, ldx, w3, M, N, K, bias, out, out_bf16, ld_out, workspace, (size_t)workspace_bytes, S(stream));
}
int nuwa_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, void* stream) {
  return nchw_f32_to_nhwc_bf16(in, out, B, C, H, W, S(stream));
}
int nuwa_nhwc_to_nchw_f32(const void* in, int in_is_bf16, float* out, int B, int C, int H, int W, void* stream) {
  return nhwc_to_nchw_f32(in, in_is_bf16, out, B, C, H, W, S(stream));
}
int nuwa_im2col_nchw_f32(const float* img, void* out, int B, int C, int H, int W, int KS, int Kpad, void* stream) {
  return im2col_nchw_f32(img, out, B, C, H, W, KS, Kpad, S(stream));
}
int nuwa_groupnorm_nhwc(const float* x, const float* w, const float* bias, float* stats_ws, void* out_bf16,
                        float* out_f32, int B, int HW, int C, int G, int leaky, void* stream) {
  return groupnorm_nhwc(x, w, bias, stats_ws, out_bf16, out_f32, B, HW, C, G, leaky, S(stream));
}
int nuwa_upsample2x_nhwc_bf16(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  return upsample2x_nhwc_bf16(in, out, B, H, W, C, S(stream));
}
int nuwa_vae_attn_prep(const float* qkv, void* out, int B, int n, int inner, void* stream) {
  return vae_attn_prep(qkv, out, B, n, inner, S(stream));
}
unsigned long long nuwa_vq_argmax_tc_workspace(int M, int Kc, int D) { return (unsigned long long)vq_argmax_tc_workspace(M, Kc, D); }
int nuwa_vq_argmax_tc(const float* x, const float* code, const float* code_sq, const void* code_bf16, const float* emax,
                      long long* out, int M, int Kc, int D, int cosine, void* workspace, unsigned long long workspace_bytes,
                      void* stream) {
  return vq_argmax_tc(x, code, code_sq, code_bf16, emax, out, M, Kc, D, cosine, workspace, (size_t)workspace_bytes, S(stream));
}
int nuwa_vq_argmax(const float* x, const float* code, const float* code_sq, long long* out, int M, int Kc, int D,
                   int cosine, void* stream) {
  return vq_argmax(x, code, code_sq, out, M, Kc, D, cosine, S(stream));
}
int nuwa_gather_rows(const float* table, const long long* idx, void* out_bf16, float* out_f32, long long M, int D,
                     void* stream) {
  return gather_rows(table, idx, out_bf16, out_f32, M, D, S(stream));
}
int nuwa_conv1x1_nhwc_to_nchw(const void* x, const float* w, const float* bias, float* out, int B, int HW, int C,
                              int Cout, void* stream) {
  return conv1x1_nhwc_to_nchw(x, w, bias, out, B, HW, C, Cout, S(stream));
}

}  // extern "C"

extern "C" void* nuwa_gemm_prof_open(void) { return gemm_prof_open(); }
extern "C" void nuwa_gemm_prof_attach(void* h) { gemm_prof_attach(h); }
extern "C" int nuwa_gemm_prof_collect(void* h, double* flops, float* ms) { return gemm_prof_collect(h, flops, ms); }
extern "C" double nuwa_gemm_prof_bytes(void* h) { return gemm_prof_bytes(h); }
extern "C" void nuwa_gemm_prof_close(void* h) { gemm_prof_close(h); }

// sizes of the parameter structs, so a foreign-language binding can verify its mirror of the layout
extern "C" void nuwa_struct_sizes(int* out3) {
  out3[0] = (int)sizeof(nuwa_ln_params);
  out3[1] = (int)sizeof(nuwa_attn_params);
  out3[2] = (int)sizeof(nuwa_embed_params);
}

// ---------------- training (backward) entry points ----------------
extern "C" {
int nuwa_gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32,
                          int ld_out, int splits, int force_bn, void* stream) {
  return gemm_bf16(A, lda, W, ldw, M, N, K, nullptr, nullptr, 0, out_f32, nullptr, ld_out, ACT_NONE, force_bn, S(stream),
                   splits < 2 ? 2 : splits);
}
int nuwa_gemm_bf16_tn_splitk(const void* At, int lda, const void* Wt, int ldw, int M, int N, int K, float* out_f32,
                             int ld_out, int splits, int force_bn, void* stream) {
  return gemm_bf16_tn_splitk(At, lda, Wt, ldw, M, N, K, out_f32, ld_out, splits < 2 ? 2 : splits, force_bn, S(stream));
}
int nuwa_bgemm(const nuwa_bgemm_params* p, void* stream) { return p ? bgemm(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_ln_bwd_grid(int rows) { return ln_bwd_grid(rows); }
int nuwa_ln_bwd(const nuwa_lnbwd_params* p, void* stream) { return p ? ln_bwd(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_reduce_partials(const float* part, int nparts, int D, float* o0, float* o1, float* o2, void* stream) {
  return reduce_partials(part, nparts, D, o0, o1, o2, S(stream));
}
int nuwa_transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int R, int C, void* stream) {
  return transpose_bf16(in, ld_in, out, ld_out, R, C, S(stream));
}
int nuwa_geglu_fwd(const void* h, void* g, long long M, int ip, void* stream) { return geglu_fwd(h, g, M, ip, S(stream)); }
int nuwa_geglu_bwd(const void* dg, const void* h, void* dh, long long M, int ip, void* stream) {
  return geglu_bwd(dg, h, dh, M, ip, S(stream));
}
int nuwa_ce_bwd(const float* logits, int ld, const long long* target, const float* gscale, void* dlogits, int ld_out,
                int rows, int V, void* stream) {
  return ce_bwd(logits, ld, target, gscale, dlogits, ld_out, rows, V, S(stream));
}
int nuwa_embed_bwd(const nuwa_embed_bwd_params* p, void* stream) { return p ? embed_bwd(*p, S(stream)) : NUWA_ERR_INVALID; }
int nuwa_rotary_bwd_to_bf16(const float* dqkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                            void* stream) {
  return rotary_bwd_to_bf16(dqkv, out, inv_freq, rows, n, H, dh, rot, S(stream));
}
int nuwa_add_rows_f32(float* dst, long long ld_dst, const float* src, long long ld_src, const int* map, int rows, int cols,
                      int accumulate, void* stream) {
  return add_rows_f32(dst, ld_dst, src, ld_src, map, rows, cols, accumulate, S(stream));
}
int nuwa_attn_bwd_rows(const nuwa_attn_rows_params* p, void* stream) {
  return p ? attn_bwd_rows(*p, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_kv_full_build(const void* k, const void* v, long long kv_bs, int kv_rs, const float* null_k, const float* null_v,
                       void* kfull, void* vfull, int B, int nk, int jp, int inner, void* stream) {
  return kv_full_build(k, v, kv_bs, kv_rs, null_k, null_v, kfull, vfull, B, nk, jp, inner, S(stream));
}
int nuwa_kv_full_split(const float* dkfull, const float* dvfull, float* dnull_k, float* dnull_v, void* dk16, void* dv16,
                       float* dk32, float* dv32, long long o_bs, int o_rs, int B, int nk, int jp, int inner, void* stream) {
  return kv_full_split(dkfull, dvfull, dnull_k, dnull_v, dk16, dv16, dk32, dv32, o_bs, o_rs, B, nk, jp, inner, S(stream));
}
int nuwa_mask_scores(float* Sc, const unsigned char* mask, int mask_bs, int B, int H, int nq, int jp, int nk, int has_null,
                     void* stream) {
  return mask_scores(Sc, mask, mask_bs, B, H, nq, jp, nk, has_null, S(stream));
}
int nuwa_attn_dense_q1(const nuwa_attn_params* p, int nk, void* stream) {
  return p ? attn_dense_q1(*p, nk, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_dense_q1_bwd(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, void* dq, long long dq_bs,
                           float* dk, float* dv, long long dkv_bs, int dkv_rs, float* dnull_k, float* dnull_v, void* stream) {
  return p ? attn_dense_q1_bwd(*p, nk, dO, do_bs, dq, dq_bs, dk, dv, dkv_bs, dkv_rs, dnull_k, dnull_v, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_dense_bwd_fused(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, int do_rs, void* Pp,
                              void* dS, int jp, float* dtalk, float out_scale, void* stream) {
  return p ? attn_dense_bwd_fused(*p, nk, dO, do_bs, do_rs, Pp, dS, jp, dtalk, out_scale, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn3dna_bwd_dq_umma(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                              void* stream) {
  return p ? attn_3dna_umma_dq(*p, dS, jp, dq, dq_bs, dq_rs, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attnx2_bwd_dq_umma(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                            void* stream) {
  return p ? attn_cross2dna_umma_dq(*p, dS, jp, dq, dq_bs, dq_rs, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn3dna_bwd_scores_umma(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* Sc, float* dPp,
                                  int jp, void* stream) {
  return p ? attn_3dna_umma_scores(*p, dO, do_bs, do_rs, Sc, dPp, jp, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn3dna_bwd_scores(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* Sc, float* dPp,
                             int jp, void* stream) {
  return p ? attn3dna_bwd_scores(*p, dO, do_bs, do_rs, Sc, dPp, jp, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn3dna_bwd_dq(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                         void* stream) {
  return p ? attn3dna_bwd_dq(*p, dS, jp, dq, dq_bs, dq_rs, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn3dna_bwd_dkdv(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, const void* dS,
                           const void* Pp, int jp, void* dk, void* dv, long long dkv_bs, int dkv_rs, void* stream) {
  return p ? attn3dna_bwd_dkdv(*p, dO, do_bs, do_rs, dS, Pp, jp, dk, dv, dkv_bs, dkv_rs, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attn_bwd_first_key(const void* q, long long q_bs, int q_rs, const void* dO, long long do_bs, int do_rs,
                            const void* dS, const void* Pp, int jp, int B, int H, int dh, int nq, float* out_k,
                            float* out_v, long long ok_bs, void* stream) {
  return attn_bwd_first_key(q, q_bs, q_rs, dO, do_bs, do_rs, dS, Pp, jp, B, H, dh, nq, out_k, out_v, ok_bs, S(stream));
}
int nuwa_attn3dna_bwd_first_key_finalize(const float* tmp_k, const float* tmp_v, const void* dO_bos, long long do_bs,
                                         void* dqkv, long long dqkv_bs, int inner, int B, void* stream) {
  return attn3dna_bwd_first_key_finalize(tmp_k, tmp_v, dO_bos, do_bs, dqkv, dqkv_bs, inner, B, S(stream));
}
int nuwa_attnx2_bwd_scores_umma(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* Sc, float* dPp,
                                int jp, void* stream) {
  return p ? attn_cross2dna_umma_scores(*p, dO, do_bs, do_rs, Sc, dPp, jp, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attnx2_bwd_scores(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* Sc, float* dPp,
                           int jp, void* stream) {
  return p ? attnx2_bwd_scores(*p, dO, do_bs, do_rs, Sc, dPp, jp, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attnx2_bwd_dq(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, void* stream) {
  return p ? attnx2_bwd_dq(*p, dS, jp, dq, dq_bs, dq_rs, S(stream)) : NUWA_ERR_INVALID;
}
int nuwa_attnx2_bwd_dkdv(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, int do_rs, const void* dS,
                         const void* Pp, int jp, const float* base_k, const float* base_v, long long base_bs, int base_rs,
                         void* dk, void* dv, long long dkv_bs, int dkv_rs, void* stream) {
  return p ? attnx2_bwd_dkdv(*p, nk, dO, do_bs, do_rs, dS, Pp, jp, base_k, base_v, base_bs, base_rs, dk, dv, dkv_bs, dkv_rs,
                             S(stream))
           : NUWA_ERR_INVALID;
}
void nuwa_struct_sizes_bwd(int* out4) {
  out4[0] = (int)sizeof(nuwa_bgemm_params);
  out4[1] = (int)sizeof(nuwa_lnbwd_params);
  out4[2] = (int)sizeof(nuwa_embed_bwd_params);
  out4[3] = (int)sizeof(nuwa_attn_rows_params);
}
}
