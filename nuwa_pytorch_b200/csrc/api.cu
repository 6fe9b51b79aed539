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
  if (!p) return NUWA_ERR_INVALID;
  if (vt_workspace != nullptr) {
    const int rc = attn_3dna_tc(*p, vt_workspace, S(stream));
    if (rc != NUWA_ERR_INVALID) return rc;  // outside the tensor-core kernel's envelope -> generic kernel
  }
  return attn_sparse3dna(*p, S(stream));
}
int nuwa_attn_dense(const nuwa_attn_params* p, void* vt_workspace, void* stream) {
  if (!p) return NUWA_ERR_INVALID;
  if (vt_workspace != nullptr && p->nq >= 8) {
    const int nk = p->jmax - (p->null_k != nullptr ? 1 : 0);
    const int rc = attn_dense_mma(*p, nk, vt_workspace, S(stream));
    if (rc != NUWA_ERR_INVALID) return rc;  // unsupported shape -> generic kernel below
  }
  return attn_dense(*p, S(stream));
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

extern "C" void nuwa_gemm_prof_enable(int on) { gemm_prof_enable(on); }
extern "C" int nuwa_gemm_prof_collect(double* flops, float* ms) { return gemm_prof_collect(flops, ms); }

// sizes of the parameter structs, so a foreign-language binding can verify its mirror of the layout
extern "C" void nuwa_struct_sizes(int* out3) {
  out3[0] = (int)sizeof(nuwa_ln_params);
  out3[1] = (int)sizeof(nuwa_attn_params);
  out3[2] = (int)sizeof(nuwa_embed_params);
}
