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

#ifdef __cplusplus
}
#endif
#endif /* NUWA_B200_H */
