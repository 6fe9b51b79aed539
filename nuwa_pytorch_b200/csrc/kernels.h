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
};

int device_sm_count();

// gemm_tcgen05.cu
int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
              const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act, int force_bn,
              cudaStream_t stream);
int conv2d_nhwc_bf16(const void* x, const void* w, int B, int Hin, int Win, int Cin, int Cout, int ksize, int stride,
                     const float* bias, const float* residual, float* out_f32, void* out_bf16, int act, int force_bn,
                     cudaStream_t stream);

}  // namespace nuwa
