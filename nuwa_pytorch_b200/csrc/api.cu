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

}  // extern "C"
