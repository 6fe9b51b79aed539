/* nuwa_b200.h -- C-ABI of libnuwa_b200.so, the sm_100a kernel library behind the NUWA / NUWASketch /
 * VQGanVAE hot paths.
 *
 * The reference (lucidrains/nuwa-pytorch) has no FFI layer: every op is an ATen call made from
 * Python.  Each entry point below therefore names the reference op(s) it replaces (file:line under
 * /root/reference/nuwa_pytorch/).  Conventions:
 *   - every pointer is a DEVICE pointer unless it says "host"; bf16 data is passed as void*,
 *   - `stream` is a cudaStream_t passed as void*,
 *   - the library never allocates, frees or synchronises; all memory is owned by the caller,
 *   - return value: 0 on success, a negative NUWA_ERR_* code otherwise (nuwa_strerror()).
 */
#ifndef NUWA_B200_H
#define NUWA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUWA_OK 0
#define NUWA_ERR_INVALID (-1)
#define NUWA_ERR_CUDA (-2)
#define NUWA_ERR_DRIVER (-3)
#define NUWA_ERR_WORKSPACE (-4)

#define NUWA_ACT_NONE 0
#define NUWA_ACT_LEAKY 1 /* LeakyReLU(0.1): vqgan_vae.py:94-95 */
#define NUWA_ACT_GLU 2   /* nn.GLU(dim=1) on packed pairs: vqgan_vae.py:217,220 */
#define NUWA_ACT_GEGLU 3 /* x * gelu(gate): nuwa_pytorch.py:255-258 */

const char* nuwa_strerror(int code);
int nuwa_abi_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
unsigned long long nuwa_launch_count(void);
/* sizeof(nuwa_ln_params), sizeof(nuwa_attn_params), sizeof(nuwa_embed_params) -- lets a binding verify its struct mirror */
void nuwa_struct_sizes(int* out3);

/* Measurement aid for bench.py's roofline.  A profiler is an object, not library state: open one, ATTACH it to the calling
 * host thread, and every launch of the tcgen05 GEMM / conv kernel made by that thread is bracketed by CUDA events on its
 * launch stream and counted.  nuwa_gemm_prof_collect (call after synchronising) returns the number of launches and their
 * summed algorithmic FLOPs and device time, then resets the counters; nuwa_gemm_prof_bytes returns the algorithmic HBM
 * bytes (every operand read once, every output written once) of the launches the last collect summed up -- the
 * denominator for the ncu DRAM-traffic comparison.  Attach NULL to detach; close destroys the events.  Threads without an
 * attached profiler record nothing, so the library holds no global mutable state. */
void* nuwa_gemm_prof_open(void);
void nuwa_gemm_prof_attach(void* prof);
int nuwa_gemm_prof_collect(void* prof, double* flops, float* ms);
double nuwa_gemm_prof_bytes(void* prof);
void nuwa_gemm_prof_close(void* prof);

/* ---- dense contraction (tcgen05) ------------------------------------------------------------
 * out[M,N] = act(A[M,K] @ W[N,K]^T + bias) + residual.   A, W bf16 with K contiguous.
 * Replaces nn.Linear at nuwa_pytorch.py:274,277 (FeedForward), :311-313 (Attention), :401-405
 * (Sparse3DNA), :783-785 (SparseCross2DNA), :1819 (to_logits), and the 1x1 nn.Conv2d at
 * vqgan_vae.py:222,238,262-263 plus VectorQuantize project_in/out (vqgan_vae.py:368-378).
 * For ACT_GLU / ACT_GEGLU the N rows of W are "pair packed": in every block of 32 rows the first 16
 * are value rows and the last 16 the matching gate rows; the output then has N/2 columns.
 * force_bn: 0 = auto, or 64/128/256 to pin the N tile. */
int nuwa_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act,
                   int force_bn, void* stream);

/* ---- implicit-GEMM convolution (tcgen05) ----------------------------------------------------
 * NHWC bf16 input x[B,Hin,Win,Cin]; weights w[Cout][KH*KW][Cin_pad] bf16 (Cin_pad = Cin rounded up
 * to 64, zero filled); output NHWC [B,H,W,Cout] (Cout/2 for GLU).  Supported (ksize,stride,pad):
 * (3,1,1) (1,1,0) (4,2,1).  Replaces nn.Conv2d at vqgan_vae.py:216,219,232,235 (3x3), :352 (4x4
 * stride 2), :353 (3x3 after the upsample). */
int nuwa_conv2d_nhwc_bf16(const void* x, const void* w, int B, int Hin, int Win, int Cin, int Cout, int ksize,
                          int stride, const float* bias, const float* residual, float* out_f32, void* out_bf16,
                          int act, int force_bn, void* stream);


/* ---- row normalisation ------------------------------------------------------------------------
 * nuwa_sandwich_ln: stage A (if y != NULL)   x_out = res_in + LayerNorm(y; post_w, post_b)
 *                                            (SandwichNorm post-norm + residual: nuwa_pytorch.py:127,1175-1180;
 *                                             reversible y1 = x1 + f(x2): reversible.py:65-66; LayerNormChan +
 *                                             residual on NHWC rows: vqgan_vae.py:140-143,286)
 *                   stage B (if pre_w != NULL) a_out = bf16(LayerNorm(x; pre_w, pre_b)) with the
 *                                            ShiftVideoTokens channel shift fused as a scatter
 *                                            (nuwa_pytorch.py:125,200-253).
 * Rows are (b, t) with t = t0 + local index, nt rows per sample; y/res_in/x_out are dense [B*nt, D] fp32.
 * a_out row of position t lives at a_out + b*a_bs + (t - a_t0)*a_rs ; a_npos = positions the buffer holds. */
typedef struct {
  const float* y;
  const float* post_w;
  const float* post_b;
  const float* res_in;
  float* x_out;
  void* x_out_bf16;
  const float* pre_w;
  const float* pre_b;
  void* a_out;
  long long a_bs;
  int a_rs, a_t0, a_npos;
  int shift, fmap;
  int t0, B, nt, D;
  float eps;
  /* CUDA-graph decode support: when t0_ptr != NULL the position is read from device memory (t0 = *t0_ptr), so one
   * captured decode step can be replayed for every token.  gather = 1 selects the gather form of the token shift:
   * the complete operand row of position t is written to the dense a_out[b*a_bs + local*a_rs] (fixed address), its
   * shifted channels being read from `shift_cache` (bf16 [B][a_npos][D], row t holds LayerNorm(x_t)[0, D/2)). */
  const int* t0_ptr;
  int gather;
  void* shift_cache;
  long long sc_bs;
} nuwa_ln_params;
int nuwa_sandwich_ln(const nuwa_ln_params* p, void* stream);

/* StableLayerNorm of (a [+ b2]): LN(v / amax(v)) -- nuwa_pytorch.py:88-95 (b2: the reversible stream sum,
 * reversible.py:142).  Writes fp32 and/or bf16. */
int nuwa_stable_ln(const float* a, const float* b2, const float* w, const float* bias, float* out_f32, void* out_bf16,
                   int rows, int D, void* stream);

/* ---- fused attention cores -------------------------------------------------------------------
 * q/k/v/o are bf16; element strides: *_bs per batch sample, *_rs per token row; heads are contiguous
 * [H][dh] inside a row.  nq query positions per sample starting at absolute position t0.
 * talk: fp32 [H][H] talking-heads matrix or NULL.  jmax = key slots per query. */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  long long q_bs, k_bs, v_bs, o_bs;
  int q_rs, k_rs, v_rs, o_rs;
  int B, nq, t0, H, dh;
  float qscale;
  const float* head_scale; /* [H] logit multiplier (exp(scale), vqgan_vae.py:275) or NULL */
  const float* talk;
  const float* null_k;     /* fp32 [H*dh] learned null key / value (nuwa_pytorch.py:306-307) or NULL */
  const float* null_v;
  const unsigned char* key_mask; /* [B][mask_bs] 1 = attend, or NULL */
  int mask_bs;
  const float* bias;       /* [H][bias_nq][bias_nk] additive logit bias (vqgan_vae.py:277) or NULL */
  int bias_nq, bias_nk;
  /* Sparse3DNA geometry (nuwa_pytorch.py:407-457) */
  int fmap, max_frames, nv, kt, kh, kw, dt, dh_, dw, causal;
  /* SparseCross2DNA geometry (nuwa_pytorch.py:789-792) */
  int ck, cdil;
  int jmax;
  int nk_dense; /* set by the library (dense tensor-core path); callers leave it 0 */
  const int* t0_ptr; /* CUDA-graph decode: if non-NULL, t0 = *t0_ptr (and, for Sparse3DNA, nv = t0) */
} nuwa_attn_params;
/* Sparse3DNA.forward core, nuwa_pytorch.py:490-608 (jmax = 1 + kt*kh*kw; k/v row 0 = bos, row 1+i = video token i) */
int nuwa_attn_sparse3dna(const nuwa_attn_params* p, void* vt_workspace, void* stream);
/* vt_workspace: unused (pass NULL); kept so the ABI of round 1 stays valid.  Generic gather kernel: any head geometry,
 * causal or centred windows, decode steps (t0_ptr). */
/* Same op on the halo-tiled tensor-core kernel (attention_3dna_halo.cu): one CTA = 4 query rows of a frame, the
 * key rows they share staged in shared memory by TMA, banded 16x16 score / PV blocks on mma.sync.  Envelope: causal
 * full pass (t0 == 0, nq == nv + 1), 16-wide token grid, H == 8, dh == 64, kh <= 3, <= 47 window keys, q|k|v in one
 * row-strided buffer; returns NUWA_ERR_INVALID outside it (nothing launched) so the caller can use the gather kernel. */
int nuwa_attn_sparse3dna_halo(const nuwa_attn_params* p, void* stream);
/* Same op on the 5th-generation tensor cores (attention_3dna_umma.cu): tcgen05.mma with TMEM accumulators.  One CTA per
 * SM walks 128-query tiles (8 dilation-spaced grid rows x 16 columns, all heads): S = Q K^T of a (head, frame offset)
 * unit is ONE UMMA against the <= 10 key rows the tile shares (TMA-staged), the band is extracted thread-per-query from
 * TMEM, softmax in fp32, probabilities parked in TMEM, talking heads in fp32 registers, and P'V runs with the A operand
 * read from tensor memory and V as MN-major B operand.  Envelope: full pass (t0 == 0, nq == nv + 1), 16-wide grid, H == 8,
 * dh == 64, kw == 3, kh <= 3, kt <= 5, column dilation 1 / 2 / 4, causal or centred window; NUWA_ERR_INVALID outside it
 * (nothing launched). */
int nuwa_attn_sparse3dna_umma(const nuwa_attn_params* p, void* stream);
/* SparseCross2DNA.forward non-bos queries, nuwa_pytorch.py:851-895, on the same tcgen05 / TMEM kernel: p->q / p->o point at
 * the first non-bos query (t0 == 1), slot 0 = the learned null key / value (fp32), unit a = context frame a, centred
 * ck x ck window with dilation cdil at the query's own grid position, context mask applied to the gathered scores.
 * Envelope: 16-wide grid, H == 8, dh == 64, ck == 3, cdil 1 / 2 / 4, 1..5 context frames, k and v rows sharing one token
 * stride; NUWA_ERR_INVALID outside it (nothing launched; nuwa_attn_cross2dna covers the rest). */
int nuwa_attn_cross2dna_umma(const nuwa_attn_params* p, void* stream);
/* Attention.forward core, nuwa_pytorch.py:339-378, and VQGanAttention core, vqgan_vae.py:275-282
 * (jmax = nk (+1 with a null key)) */
int nuwa_attn_dense(const nuwa_attn_params* p, void* vt_workspace, void* stream);
/* vt_workspace: NULL, or B*H*dh*roundup(nk,64) bf16 elements of scratch; when given and nk <= 256, dh in {32,64},
 * H <= 8, nq >= 8 the tensor-core (mma.sync) variant runs, otherwise the generic CUDA-core kernel. */
/* Same op on the probability-resident kernel (attention_dense_pres.cu): 32-query tiles, the fp16 probability slab of
 * all 8 heads kept in shared memory, K / V streamed once per tile by TMA, talking-heads mix on the tensor cores.
 * Envelope: H == 8, dh == 64, 1 <= keys <= 256 (p->jmax minus the null slot), no bias / head_scale; returns
 * NUWA_ERR_INVALID outside it (nothing launched). */
int nuwa_attn_dense_pres(const nuwa_attn_params* p, void* stream);
/* SparseCross2DNA.forward non-bos queries, nuwa_pytorch.py:851-895 (jmax = 1 + frames*ck*ck; t0 >= 1) */
int nuwa_attn_cross2dna(const nuwa_attn_params* p, void* stream);

/* ---- token level ----------------------------------------------------------------------------- */
typedef struct {
  float* out;             /* [B*nt, D] fp32 */
  const long long* idx;   /* token ids, row b at idx + b*idx_bs */
  long long idx_bs;
  const float* table;     /* [V, D] */
  const float* bos;       /* [D] or NULL */
  const float* ax1;       /* axial tables (NULL = axis dropped), position p -> (p/(d2*d3), (p/d3)%d2, p%d3) */
  const float* ax2;
  const float* ax3;
  int d2, d3;
  int has_bos, t0, B, nt, D;
  const int* t0_ptr; /* CUDA-graph decode: if non-NULL, t0 = *t0_ptr */
} nuwa_embed_params;
/* Embedding + AxialPositionalEmbedding + bos: nuwa_pytorch.py:1659-1709,1879-1881,1940-1944 */
int nuwa_embed_tokens(const nuwa_embed_params* p, void* stream);
/* rotary embedding of q,k AND v (nuwa_pytorch.py:132-153,333-335): qkv fp32 [rows,3*H*dh] -> bf16 */
int nuwa_rotary_to_bf16(const float* qkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                        void* stream);
/* F.cross_entropy mean over rows (nuwa_pytorch.py:1963); row_loss: workspace [rows] */
int nuwa_cross_entropy_mean(const float* logits, int ld, const long long* target, float* row_loss, float* out, int rows,
                            int V, void* stream);
/* generate() sampling step (nuwa_pytorch.py:1901-1906,1713-1719,55-66): guidance mix, top-k, gumbel argmax.
 * uncond may be NULL (cond_scale == 1); guided_out (optional) receives the mixed logits [B,V]. */
int nuwa_sample_topk_gumbel(const float* cond, const float* uncond, const float* noise, long long* out,
                            float* guided_out, int B, int V, int k, float cond_scale, float temperature, void* stream);
/* Same, for a captured decode step: step = *step_ptr (device); noise row = noise + step*B*V; the sampled id of sample b
 * is written to out[b*out_bs + step] (the token sequence buffer the next step's embedding reads). */
int nuwa_sample_topk_gumbel_at(const float* cond, const float* uncond, const float* noise, long long* out,
                               long long out_bs, const int* step_ptr, int B, int V, int k, float cond_scale,
                               float temperature, void* stream);
/* Decode-step helpers: cache[b][*t_ptr][:] = row[b][:] (bf16, `width` elements; appends the new token's q|k|v to the KV
 * cache) and *t_ptr += 1. */
int nuwa_cache_append(const void* row, void* cache, long long cache_bs, int width, int B, const int* t_ptr, void* stream);
int nuwa_step_increment(int* t_ptr, void* stream);

/* ---- persistent decode step (generate()) -------------------------------------------------------
 * ONE cooperative kernel launch runs a whole Transformer / ReversibleTransformer stack for the single new position
 * *t_ptr of every sample (nuwa_pytorch.py:1168-1182, :1289-1295 and reversible.py:61-68,132-142 restricted to one
 * token, over SandwichNorm :112-128, ShiftVideoTokens :200-253, Sparse3DNA :459-613, Attention :315-379, FeedForward
 * :255-286, StableLayerNorm :88-95; optionally to_logits :1819).  It replaces the ~15 launches per layer of the
 * per-kernel decode path; see csrc/decode_stack.cu.  All pointers are device pointers; `subs` is a DEVICE array. */
#define NUWA_DEC_3DNA 0
#define NUWA_DEC_CROSS 1
#define NUWA_DEC_FF 2
typedef struct {
  int kind;           /* NUWA_DEC_* */
  int shift;          /* ShiftVideoTokens (space) wraps the block */
  int read, write;    /* residual stream read / updated: 0,0 (plain) or f: 1,0 / g: 0,1 (reversible.py:65-68) */
  const float* pre_w; /* SandwichNorm prenorm / postnorm affine parameters, fp32 [D] */
  const float* pre_b;
  const float* post_w;
  const float* post_b;
  const void* w_a;    /* bf16: 3DNA to_q|to_kv [3*inner][D]; cross to_q [inner][D]; FF net.0 pair packed [2*ip][D] */
  const void* w_b;    /* bf16: to_out [D][inner]; FF net.3 [D][ip] (zero padded K) */
  const float* b_out; /* Sparse3DNA to_out bias [D] or NULL */
  const float* talk;  /* talking-heads matrix fp32 [H][H] (attention kinds) */
  const float* null_k;/* cross attention learned null key / value, fp32 [H*dh] */
  const float* null_v;
  void* cache;        /* 3DNA: bf16 q|k|v cache [B][npos][3*inner] (row t is WRITTEN); cross: bf16 k|v of the context
                         [B][nk][2*inner] */
  void* shift_cache;  /* bf16 [B][npos][D]: pre-norm rows of earlier positions (row t is written) or NULL */
  int ip;             /* FF: padded inner width */
  int kt, kh, kw;     /* Sparse3DNA window (nuwa_pytorch.py:407-422); kernel and dilation are per-layer settings */
  int dt, dh_, dw;    /* (sparse_3dna_kernel_size / sparse_3dna_dilation are cast to one value per layer, :1097-1100) */
  int reserved;
} nuwa_decode_sub;
typedef struct {
  const nuwa_decode_sub* subs;
  int nsubs;
  int B, D, H, dh, npos, reversible;
  int fmap, max_frames, causal; /* Sparse3DNA token grid shared by all 3DNA sub-blocks */
  int j3max;                    /* max over the 3DNA sub-blocks of 1 + kt*kh*kw */
  int nk;                        /* context tokens of the cross attention (0 if none) */
  const unsigned char* key_mask; /* [B][mask_bs], 1 = attend (context_mask), or NULL */
  int mask_bs;
  const int* t_ptr;              /* position of the new token (device scalar) */
  const float* x_in;             /* [B][D] fp32: embedded token (or the previous sweep's output) */
  const float* norm_w;           /* final StableLayerNorm */
  const float* norm_b;
  float* out_f32;                /* [B][D] normalised output, fp32 and / or bf16 (either may be NULL) */
  void* out_bf16;
  const void* w_logits;          /* bf16 [V][D] or NULL */
  int V;
  float* logits;                 /* [B][V] fp32 */
  /* scratch, all owned by the caller */
  float* y;                      /* [B][D] fp32 */
  void* act;                     /* bf16 [B][kmax] */
  void* actq;                    /* bf16 [B][H*dh] */
  float* scores;                 /* fp32 [B][H][nk+1] (cross attention) */
  unsigned int* barrier;         /* 2 words, zero before the FIRST launch; the kernel leaves them zero */
  int kmax;                      /* max(D, H*dh, every ip) */
  int jmax;                      /* set by the library */
  int split_small, split_ff, split_logits; /* K slices (1 or 7 = worker warps per CTA; 0 = default) per 16-row tile of the
                                  D- / inner-wide products, the FF-out product and the logits product */
  int max_ctas;                  /* 0 = one CTA per SM */
  /* measurement aids (tools/decode_phase_profile.py): prof != NULL makes CTA prof_cta record (tag, clock64) pairs at
   * every phase boundary (room for 2 * (12 * nsubs + 8) values); debug_flags bit 0/1/2 skip the norms / skinny
   * products / attention (timing experiments only: the result is then meaningless) */
  long long* prof;
  int prof_cta;
  int debug_flags;
} nuwa_decode_params;
/* cooperative = 1: cudaLaunchCooperativeKernel (co-residency guaranteed by the driver); 0: plain launch with
 * grid <= SM count.  Returns NUWA_ERR_INVALID for shapes outside the kernel's envelope (B > 16, D > 1024, ...). */
int nuwa_decode_stack(const nuwa_decode_params* p, int cooperative, void* stream);
/* sizeof(nuwa_decode_sub), sizeof(nuwa_decode_params) */
void nuwa_struct_sizes_decode(int* out2);

/* ---- VQGanVAE support kernels ---------------------------------------------------------------- */
int nuwa_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, void* stream);
int nuwa_nhwc_to_nchw_f32(const void* in, int in_is_bf16, float* out, int B, int C, int H, int W, void* stream);
/* im2col of the first conv (vqgan_vae.py:365): K index (kh*KS+kw)*C + c, zero padded to Kpad */
int nuwa_im2col_nchw_f32(const float* img, void* out, int B, int C, int H, int W, int KS, int Kpad, void* stream);
/* nn.GroupNorm(G, C) on NHWC fp32 (+LeakyReLU 0.1) (vqgan_vae.py:218,221,233,236); stats_ws: [B*G*2] floats */
int nuwa_groupnorm_nhwc(const float* x, const float* w, const float* bias, float* stats_ws, void* out_bf16,
                        float* out_f32, int B, int HW, int C, int G, int leaky, void* stream);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=False) (vqgan_vae.py:353) */
int nuwa_upsample2x_nhwc_bf16(const void* in, void* out, int B, int H, int W, int C, void* stream);
/* VQGanAttention q,k l2norm over the spatial axis (vqgan_vae.py:271-273): qkv fp32 [B,n,3*inner] -> bf16 */
int nuwa_vae_attn_prep(const float* qkv, void* out, int B, int n, int inner, void* stream);
/* VectorQuantize codebook arg-max (call site vqgan_vae.py:435). code: cosine -> pre-normalised codebook,
 * euclid -> raw codebook + code_sq[Kc] = |e|^2.  out: int64 [M], first maximum wins. */
int nuwa_vq_argmax(const float* x, const float* code, const float* code_sq, long long* out, int M, int Kc, int D,
                   int cosine, void* stream);
/* Same arg-max with the M x Kc x D contraction on the tensor cores and an exact fp32 re-score of every code inside the
 * provable bf16 error band of the row maximum (csrc/vae_ops.cu: vq_argmax_tc), so the result is the fp32 arg-max.
 * code_bf16: bf16 copy of `code`; emax: DEVICE scalar holding max_j |code_j| (read by the kernel, so packing the
 * codebook needs no host synchronisation); workspace:
 * nuwa_vq_argmax_tc_workspace(M, Kc, D) bytes of device memory, 16-byte aligned.  Envelope: D % 8 == 0, D <= 1024,
 * Kc % 4 == 0; NUWA_ERR_INVALID outside it. */
unsigned long long nuwa_vq_argmax_tc_workspace(int M, int Kc, int D);
int nuwa_vq_argmax_tc(const float* x, const float* code, const float* code_sq, const void* code_bf16, const float* emax,
                      long long* out, int M, int Kc, int D, int cosine, void* workspace, unsigned long long workspace_bytes,
                      void* stream);
/* fp32-faithful nn.Linear on the tensor cores: VectorQuantize.project_in (vqgan_vae.py:368-378, third-party
 * vector_quantize_pytorch `project_in`) feeds the codebook arg-max, so its result must be the fp32 reference's, not a
 * bf16-operand approximation.  Every fp32 operand is split exactly into three bf16 terms (x = x0 + x1 + x2) and the six
 * products with i + j <= 2 are accumulated in fp32 by six tcgen05 GEMM launches (csrc/vae_ops.cu).
 * nuwa_split3_f32_bf16: x[rows][K] fp32 (row stride ld) -> out[rows][3K] bf16 = [x0 | x1 | x2]; used once per weight
 * version for W.  nuwa_linear_f32x3: out[M,N] fp32 (and, if out_bf16 != NULL, its bf16 rounding) = x[M,K] @ W^T + bias with w3 = split of W[N][K]; workspace =
 * nuwa_linear_f32x3_workspace(M, K) bytes.  K % 8 == 0. */
int nuwa_split3_f32_bf16(const float* x, long long ld, void* out, long long rows, int K, void* stream);
unsigned long long nuwa_linear_f32x3_workspace(int M, int K);
int nuwa_linear_f32x3(const float* x, long long ldx, const void* w3, int M, int N, int K, const float* bias, float* out,
                      void* out_bf16, int ld_out, void* workspace, unsigned long long workspace_bytes, void* stream);
/* F.embedding / codebook[indices] (vqgan_vae.py:447, nuwa_pytorch.py:1910) */
int nuwa_gather_rows(const float* table, const long long* idx, void* out_bf16, float* out_f32, long long M, int D,
                     void* stream);
/* final Conv2d(dim, channels, 1) (vqgan_vae.py:366): NHWC bf16 -> NCHW fp32, Cout <= 8 */
int nuwa_conv1x1_nhwc_to_nchw(const void* x, const float* w, const float* bias, float* out, int B, int HW, int C,
                              int Cout, void* stream);

/* =================================================================================================
 * Training (backward) entry points -- the gradient of `loss = nuwa(text=, video=, return_loss=True); loss.backward()`
 * (nuwa_pytorch.py:1917-1964; the reference obtains it from ATen autograd, reversible.py:70-123 for the reversible
 * stacks).  Same conventions as above; gradient accumulators are fp32 and are ADDED to unless stated.
 * ================================================================================================= */

/* out_f32[M,N] += A[M,K] @ W[N,K]^T with the contraction split over up to `splits` work items per output tile (fp32
 * atomics).  The weight-gradient GEMMs dW = dY^T X have a small output and a long contraction (K = tokens). */
int nuwa_gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out_f32,
                          int ld_out, int splits, int force_bn, void* stream);
/* out_f32[M,N] += At^T @ Wt with BOTH operands stored contraction-major: At is [K][M] (row stride lda), Wt is [K][N] (row
 * stride ldw), K = tokens.  This is dW = dY^T X (the weight gradient of nn.Linear under autograd, nuwa_pytorch.py:274,277,
 * 311-313,401-405,1819) on dY and X exactly as the passes hold them: the tcgen05 instruction reads both operands MN-major
 * from the 128-byte-swizzled tiles TMA drops (instruction-descriptor a_major = b_major = 1), no transposed copies. */
int nuwa_gemm_bf16_tn_splitk(const void* At, int lda, const void* Wt, int ldw, int M, int N, int K, float* out_f32,
                             int ld_out, int splits, int force_bn, void* stream);

/* Batched small GEMM on mma.sync (attention backward products), see csrc/bgemm.cu.
 *   C[i1][i2] (M x N) = alpha * A(M x K) * B(K x N)  (+ C if accumulate; fp32 C only)
 *   A(m,k) = a_trans ? A[k*lda + m] : A[m*lda + k] ;  B(k,n) = b_trans ? B[k*ldb + n] : B[n*ldb + k]   (bf16)
 * batch offsets (elements): i1*x_s1 + i2*x_s2.  lda/ldb and the batch strides must be multiples of 8. */
typedef struct {
  const void* A;
  const void* B;
  void* C;
  int M, N, K;
  int a_trans, b_trans;
  long long lda, ldb, ldc;
  int batch1, batch2;
  long long a_s1, a_s2, b_s1, b_s2, c_s1, c_s2;
  float alpha;
  int c_bf16, accumulate;
} nuwa_bgemm_params;
int nuwa_bgemm(const nuwa_bgemm_params* p, void* stream);

/* LayerNorm / StableLayerNorm backward of rows (b, t), nt rows per sample (nn.LayerNorm in SandwichNorm,
 * nuwa_pytorch.py:112-128; StableLayerNorm :88-95).  dout is fp32 or bf16 (exactly one non-NULL).  unshift = 1: the
 * normalised row went through ShiftVideoTokens (:200-253) before its consumer, so the upstream gradient of the first /
 * second channel quarter is read from the row one grid-row below / one column right (zero at the borders).
 * x (+ x2): the forward input of the norm; stable = 1: LN(x / amax(x)).  Outputs: dx as bf16 and/or fp32 (dx_f32 and
 * dx2_f32 are accumulated into when accumulate = 1).  The affine-parameter gradients are reduced per CTA and added to
 * dw / db / dcol with fp32 atomics. */
typedef struct {
  int rows, nt, D;
  const float* dout_f32;
  const void* dout_bf16;
  int unshift, fmap;
  const float* x;
  const float* x2;
  int stable;
  const float* w;
  float eps;
  void* dx_bf16;
  float* dx_f32;
  float* dx2_f32;
  int accumulate;
  float* dw;   /* [D] += sum_rows dout * xhat   (NULL = skip) */
  float* db;   /* [D] += sum_rows dout */
  float* dcol; /* [D] += sum_rows dx            (bias gradient of the layer that produced x) */
} nuwa_lnbwd_params;
int nuwa_ln_bwd_grid(int rows);
int nuwa_ln_bwd(const nuwa_lnbwd_params* p, void* stream);
/* o0[c] += sum_p part[p][0][c], o1 <- [1], o2 <- [2]  (NULL outputs skipped) */
int nuwa_reduce_partials(const float* part, int nparts, int D, float* o0, float* o1, float* o2, void* stream);

/* bf16 [R][ld_in] -> [C][ld_out] */
int nuwa_transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int R, int C, void* stream);

/* GEGLU (nuwa_pytorch.py:255-258) on the pair-packed pre-activation h [M][2*ip] (see NUWA_ACT_GEGLU): g [M][ip] and its
 * backward dh [M][2*ip] from dg [M][ip] */
int nuwa_geglu_fwd(const void* h, void* g, long long M, int ip, void* stream);
int nuwa_geglu_bwd(const void* dg, const void* h, void* dh, long long M, int ip, void* stream);

/* d mean-cross-entropy / d logits (F.cross_entropy, nuwa_pytorch.py:1963): bf16 dlogits = (softmax - onehot) * (*gscale) / rows;
 * gscale: device scalar (the incoming grad_output) or NULL = 1 */
int nuwa_ce_bwd(const float* logits, int ld, const long long* target, const float* gscale, void* dlogits, int ld_out,
                int rows, int V, void* stream);

/* Embedding (+frac_gradient, nuwa_pytorch.py:1659-1671), AxialPositionalEmbedding (:1693-1709) and bos backward */
typedef struct {
  const float* dx;        /* [B*nt][D] */
  const long long* idx;
  long long idx_bs;
  float* dtable;          /* [V][D] += frac * dx */
  float* dbos;            /* [D] or NULL */
  float* dax1;
  float* dax2;
  float* dax3;
  int d2, d3;
  int has_bos, B, nt, D;
  float frac;
} nuwa_embed_bwd_params;
int nuwa_embed_bwd(const nuwa_embed_bwd_params* p, void* stream);

/* inverse rotation of the gradient of the rotated q|k|v (nuwa_pytorch.py:149-153): fp32 in, bf16 out */
int nuwa_rotary_bwd_to_bf16(const float* dqkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                            void* stream);

/* dst[map[r]][0:cols] (+)= src[r][0:cols] (map NULL = identity; negative = skip): un-packs padded / pair-packed gradients */
int nuwa_add_rows_f32(float* dst, long long ld_dst, const float* src, long long ld_src, const int* map, int rows, int cols,
                      int accumulate, void* stream);

/* ---- attention backward --------------------------------------------------------------------
 * Shared row kernel: softmax + talking-heads forward/backward on materialised logits.
 *   S, dPp fp32 [B][H][nq][jp] (logits incl. scale and masks; dPp = dO V^T) -> Pp (= P', bf16), dS (bf16, x out_scale),
 *   dtalk[H][H] += sum dPp[g] P[h].   J valid slots per row, pad slots of the outputs are zeroed. */
typedef struct {
  const float* S;
  const float* dPp;
  void* Pp;
  void* dS;
  const float* talk; /* [H][H] or NULL (identity) */
  float* dtalk;      /* [H][H] or NULL */
  int B, H, nq, J, jp;
  float out_scale;
  const unsigned char* key_mask; /* dense attention: [B][mask_bs], 0 = key (slot - has_null) is masked; or NULL */
  int mask_bs, has_null;
} nuwa_attn_rows_params;
int nuwa_attn_bwd_rows(const nuwa_attn_rows_params* p, void* stream);
/* Dense attention with ONE query per sample and no talking heads: the bos query of SparseCross2DNA.forward over
 * [null key | every context token] (nuwa_pytorch.py:828-844) in full passes, forward and backward (under autograd).  One CTA
 * per (sample, head).  p->q / p->o: the single row of every sample (q_bs / o_bs), p->nq == 1, dh == 64, p->talk == NULL.
 * Backward: dO / dq one row per sample; dk, dv fp32 [B][nk] rows, WRITTEN; dnull_k / dnull_v fp32 [H*64], ADDED to.
 * NUWA_ERR_INVALID outside the envelope (nothing launched; nuwa_attn_dense and the batched-GEMM backward cover the rest). */
int nuwa_attn_dense_q1(const nuwa_attn_params* p, int nk, void* stream);
int nuwa_attn_dense_q1_bwd(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, void* dq, long long dq_bs,
                           float* dk, float* dv, long long dkv_bs, int dkv_rs, float* dnull_k, float* dnull_v, void* stream);
/* Dense attention backward (Attention core, nuwa_pytorch.py:339-378, under autograd), probability stage fused: per tile of
 * 16 queries x 8 heads the logits S = Q K^T and dP' = dO V^T are recomputed on the tensor cores from TMA-staged K / V
 * chunks and stay in shared memory; the softmax / talking-heads backward runs from there and writes P' (operand of
 * dV = P'^T dO) and dS * out_scale (operand of dQ = dS K, dK = dS^T Q) as bf16 [B][8][nq][jp] (slot 0 = null key when
 * p->null_k is set, slot has_null + j = key j, zero padding up to jp) and adds dW_talk to dtalk (fp32 [8][8], may be NULL).
 * Replaces nuwa_bgemm (S) + nuwa_bgemm (dP') + nuwa_attn_bwd_rows and their fp32 [B][8][nq][jp] round trips through HBM.
 * p: q / k / v pointers + strides, nq, B, qscale, talk, null_k / null_v (fp32, exact null logit as in the forward kernel),
 * key_mask; dO: bf16 [B][nq] rows.  Envelope: H == 8, dh == 64, 1 <= nk <= 256; NUWA_ERR_INVALID outside it. */
int nuwa_attn_dense_bwd_fused(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, int do_rs, void* Pp,
                              void* dS, int jp, float* dtalk, float out_scale, void* stream);
/* dense attention (nuwa_pytorch.py:339-378): K / V with the learned null slot prepended, zero padded to jp rows */
int nuwa_kv_full_build(const void* k, const void* v, long long kv_bs, int kv_rs, const float* null_k, const float* null_v,
                       void* kfull, void* vfull, int B, int nk, int jp, int inner, void* stream);
int nuwa_kv_full_split(const float* dkfull, const float* dvfull, float* dnull_k, float* dnull_v, void* dk16, void* dv16,
                       float* dk32, float* dv32, long long o_bs, int o_rs, int B, int nk, int jp, int inner, void* stream);
int nuwa_mask_scores(float* S, const unsigned char* mask, int mask_bs, int B, int H, int nq, int jp, int nk, int has_null,
                     void* stream);
/* Sparse3DNA (nuwa_pytorch.py:490-608): p describes the NON-bos queries (p.t0 >= 1, p.q = first such row) */
int nuwa_attn3dna_bwd_scores(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp,
                             int jp, void* stream);
/* The same two tensors from the tcgen05 / TMEM kernel of nuwa_attn_sparse3dna_umma run in scores mode, twice (logits:
 * Q = q, K = k; dP' = dO V^T: Q = dO from its own buffer, K = v): one UMMA per (head, frame offset) of a 128-query tile
 * instead of 46 gathered key rows per query.  Same envelope as nuwa_attn_sparse3dna_umma (full pass: p->nq == p->nv,
 * p->t0 == 1, q|k|v rows in one buffer); NUWA_ERR_INVALID outside it (nothing launched). */
int nuwa_attn3dna_bwd_scores_umma(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp,
                                  int jp, void* stream);
/* nuwa_attn3dna_bwd_dq / nuwa_attnx2_bwd_dq on the tcgen05 / TMEM kernel in PV mode: dS (bf16, slot order) takes the place of
 * the probabilities and V := K, so the PV stage of the forward kernel (A operand from TMEM, K tiles MN-major by TMA) yields
 * dq = sum_j dS[j] k_j.  Kernel height 3, jp % 8 == 0, otherwise the envelopes of the forward entries; NUWA_ERR_INVALID
 * outside (nothing launched). */
int nuwa_attn3dna_bwd_dq_umma(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                              void* stream);
int nuwa_attnx2_bwd_dq_umma(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                            void* stream);
int nuwa_attn3dna_bwd_dq(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                         void* stream);
int nuwa_attn3dna_bwd_dkdv(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, const void* dS,
                           const void* Pp, int jp, void* dk, void* dv, long long dkv_bs, int dkv_rs, void* stream);
/* SparseCross2DNA non-bos queries (nuwa_pytorch.py:851-895): same three passes; nk = context tokens; base_k / base_v
 * (fp32 [B][nk] rows with strides base_bs / base_rs, or NULL) is added to the key gradients (the dense bos query's part) */
/* nuwa_attnx2_bwd_scores on the tcgen05 / TMEM kernel in scores mode (envelope of nuwa_attn_cross2dna_umma; NUWA_ERR_INVALID
 * outside it, nothing launched) */
int nuwa_attnx2_bwd_scores_umma(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp,
                                int jp, void* stream);
int nuwa_attnx2_bwd_scores(const nuwa_attn_params* p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp,
                           int jp, void* stream);
int nuwa_attnx2_bwd_dq(const nuwa_attn_params* p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, void* stream);
int nuwa_attnx2_bwd_dkdv(const nuwa_attn_params* p, int nk, const void* dO, long long do_bs, int do_rs, const void* dS,
                         const void* Pp, int jp, const float* base_k, const float* base_v, long long base_bs, int base_rs,
                         void* dk, void* dv, long long dkv_bs, int dkv_rs, void* stream);
/* slot 0 (bos key / null key): out_k[b] += sum_q dS[b,h,q,0] q[b,q] ; out_v[b] += sum_q Pp[b,h,q,0] dO[b,q] */
int nuwa_attn_bwd_first_key(const void* q, long long q_bs, int q_rs, const void* dO, long long do_bs, int do_rs,
                            const void* dS, const void* Pp, int jp, int B, int H, int dh, int nq, float* out_k,
                            float* out_v, long long ok_bs, void* stream);
int nuwa_attn3dna_bwd_first_key_finalize(const float* tmp_k, const float* tmp_v, const void* dO_bos, long long do_bs,
                                         void* dqkv, long long dqkv_bs, int inner, int B, void* stream);
/* ---- trainer step: clip_grad_norm_ + AdamW + zero_grad over flat fp32 buffers ---------------------------------
 * Replaces train_nuwa.py:256-258 (torch.nn.utils.clip_grad_norm_; optim.step(); optim.zero_grad()) with the optimizer of
 * optimizer.py:11-31 (AdamW, no weight decay on parameters with ndim < 2).  See csrc/optim.cu. */
/* out[0] (+)= sum x[i]^2, deterministic two-stage reduction; partials: nparts floats of scratch (grid = nparts CTAs) */
int nuwa_sqnorm_f32(const float* x, long long n, float* partials, int nparts, float* out, int accumulate, void* stream);
typedef struct {
  long long offset;   /* first element of the chunk in the flat buffers (multiple of 4) */
  int len;            /* elements; a chunk never crosses a parameter boundary */
  int weight_decay;   /* 1: decoupled weight decay applies (parameter ndim >= 2) */
} nuwa_opt_chunk;
typedef struct {
  float* p;           /* parameters   (flat fp32, layout of train.GradStore) */
  float* g;           /* gradients    (same layout; zeroed afterwards when zero_grad = 1) */
  float* m;           /* exp_avg */
  float* v;           /* exp_avg_sq */
  const nuwa_opt_chunk* chunks; /* DEVICE array, one CTA per chunk */
  int nchunks;
  float lr, beta1, beta2, eps, weight_decay;
  float max_norm;      /* clip_grad_norm_ threshold; <= 0 or sqnorm == NULL: no clipping */
  float grad_scale;    /* gradients are multiplied by this first (1 / grad_accum_every) */
  const float* sqnorm; /* device scalar: sum of squares of the UNscaled gradient buffer (nuwa_sqnorm_f32) */
  int step;            /* 1-based optimizer step, used when step_ptr == NULL */
  const int* step_ptr; /* VQGanVAE.forward(img, return_loss=True) without the GAN / perceptual terms (use_vgg_and_gan=False): the reconstruction
 * loss F.l1_loss (default) or F.mse_loss (l2_recon_loss=True) of vqgan_vae.py:340,502-512.  out[0] = mean |a-b| or
 * mean (a-b)^2 over n fp32 elements; partials: nparts floats of scratch (two deterministic reduction stages). */
int nuwa_recon_loss_f32(const float* a, const float* b, long long n, int l2, float* partials, int nparts, float* out,
                        void* stream);
/* device-side step counter (CUDA-graph replay) or NULL */
  int zero_grad;
} nuwa_adamw_params;
int nuwa_adamw_step(const nuwa_adamw_params* a, void* stream);
/* sizeof(nuwa_opt_chunk), sizeof(nuwa_adamw_params) */
void nuwa_struct_sizes_optim(int* out2);

/* sizeof of the five backward parameter structs (bgemm, lnbwd, embed_bwd, attn_rows, -) for binding checks */
void nuwa_struct_sizes_bwd(int* out4);

#ifdef __cplusplus
}
#endif
#endif /* NUWA_B200_H */
