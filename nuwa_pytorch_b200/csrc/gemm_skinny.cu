// Skinny-M GEMM for the decode step of generate(): out[M,N] = act(A[M,K] W[N,K]^T + bias) + residual with
// M <= 32 (one row per sample).  Such a product streams every weight exactly once and does 2*M FLOPs per
// weight byte pair: it is HBM/L2-bandwidth and launch-latency bound, a 128-row tensor-core tile would idle
// 94 % of its rows and occupy only N/64 CTAs.  Here every warp owns one output column (or one value/gate
// pair), the 32 lanes split K with 16-byte loads, A sits in shared memory, and N/8 CTAs cover the chip.
// Same operand conventions and fused epilogues as gemm_tcgen05_kernel (incl. pair-packed GLU/GEGLU rows).
#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int SK_WARPS = 8;
static constexpr int SK_MR = 8;  // rows accumulated per pass

__device__ __forceinline__ float dot8(const uint4& a, const uint4& w) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  return a0.x * w0.x + a0.y * w0.y + a1.x * w1.x + a1.y * w1.y + a2.x * w2.x + a2.y * w2.y + a3.x * w3.x + a3.y * w3.y;
}

__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, int M, int N, int K,
                   const float* __restrict__ bias, const float* __restrict__ residual, int ld_res,
                   float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int ld_out, int act) {
  extern __shared__ __align__(16) uint8_t smem_sk[];
  bf16* As = reinterpret_cast<bf16*>(smem_sk);  // [M][Kp], Kp = K rounded up to 8
  const int Kp = (K + 7) & ~7;
  const int k8 = Kp / 8;
  // issue this warp's first weight loads BEFORE staging A, so the two DRAM/L2 latencies overlap
  const bool pair_ = (act == ACT_GLU || act == ACT_GEGLU);
  const int n_out_ = pair_ ? N / 2 : N;
  const int col_ = blockIdx.x * SK_WARPS + (threadIdx.x >> 5);
  uint4 pre_w = make_uint4(0u, 0u, 0u, 0u), pre_g = pre_w;
  {
    const int c = (threadIdx.x & 31) * 8;
    if (col_ < n_out_ && c + 8 <= K) {
      const int wr = pair_ ? (col_ / 16) * 32 + (col_ % 16) : col_;
      pre_w = __ldg(reinterpret_cast<const uint4*>(W + (long long)wr * ldw + c));
      if (pair_) pre_g = __ldg(reinterpret_cast<const uint4*>(W + (long long)(wr + 16) * ldw + c));
    }
  }
  for (int i = threadIdx.x; i < M * k8; i += blockDim.x) {
    const int r = i / k8, c = (i - r * k8) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c + 8 <= K) {
      v = *reinterpret_cast<const uint4*>(A + (long long)r * lda + c);
    } else {  // K tail: element-wise, zero padded
      bf16 tmp[8];
      for (int e = 0; e < 8; ++e) tmp[e] = (c + e < K) ? A[(long long)r * lda + c + e] : __float2bfloat16(0.f);
      v = *reinterpret_cast<uint4*>(tmp);
    }
    *reinterpret_cast<uint4*>(As + (long long)r * Kp + c) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool pair = (act == ACT_GLU || act == ACT_GEGLU);
  const int n_out = pair ? N / 2 : N;
  const int col = blockIdx.x * SK_WARPS + warp;  // output column
  if (col >= n_out) return;
  // weight rows feeding this output column
  const int wrow = pair ? (col / 16) * 32 + (col % 16) : col;
  const bf16* w0 = W + (long long)wrow * ldw;
  const bf16* w1 = pair ? w0 + (long long)16 * ldw : nullptr;
  for (int m0 = 0; m0 < M; m0 += SK_MR) {
    float acc[SK_MR], accg[SK_MR];
#pragma unroll
    for (int r = 0; r < SK_MR; ++r) acc[r] = accg[r] = 0.f;
    for (int c = lane * 8; c < Kp; c += 256) {
      uint4 wv = make_uint4(0u, 0u, 0u, 0u), wg = wv;
      if (c == lane * 8 && c + 8 <= K) {
        wv = pre_w;  // prefetched above
        wg = pre_g;
      } else if (c + 8 <= K) {
        wv = __ldg(reinterpret_cast<const uint4*>(w0 + c));
        if (pair) wg = __ldg(reinterpret_cast<const uint4*>(w1 + c));
      } else {
        bf16 t0[8], t1[8];
        for (int e = 0; e < 8; ++e) {
          t0[e] = (c + e < K) ? w0[c + e] : __float2bfloat16(0.f);
          t1[e] = (pair && c + e < K) ? w1[c + e] : __float2bfloat16(0.f);
        }
        wv = *reinterpret_cast<uint4*>(t0);
        wg = *reinterpret_cast<uint4*>(t1);
      }
#pragma unroll
      for (int r = 0; r < SK_MR; ++r)
        if (m0 + r < M) {
          const uint4 av = *reinterpret_cast<const uint4*>(As + (long long)(m0 + r) * Kp + c);
          acc[r] += dot8(av, wv);
          if (pair) accg[r] += dot8(av, wg);
        }
    }
#pragma unroll
    for (int r = 0; r < SK_MR; ++r) {
      if (m0 + r >= M) continue;  // warp-uniform
      float v = warp_sum(acc[r]);
      float g = pair ? warp_sum(accg[r]) : 0.f;
      if (lane == 0) {
        const long long m = m0 + r;
        if (pair) {
          if (bias != nullptr) { v += bias[wrow]; g += bias[wrow + 16]; }
          v = (act == ACT_GLU) ? v * sigmoid_f(g) : v * gelu_erf(g);
        } else {
          if (bias != nullptr) v += bias[col];
          if (act == ACT_LEAKY) v = leaky01(v);
        }
        if (residual != nullptr) v += residual[m * ld_res + col];
        if (out_f32 != nullptr) out_f32[m * ld_out + col] = v;
        if (out_bf16 != nullptr) out_bf16[m * ld_out + col] = __float2bfloat16(v);
      }
    }
  }
}

int gemm_skinny(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act,
                cudaStream_t stream) {
  if (M <= 0 || M > 32 || N <= 0 || K <= 0) return NUWA_ERR_INVALID;
  if ((lda % 8) || (ldw % 8)) return NUWA_ERR_INVALID;
  const bool pair = (act == ACT_GLU || act == ACT_GEGLU);
  if (pair && (N % 32)) return NUWA_ERR_INVALID;
  const int Kp = (K + 7) & ~7;
  const size_t smem = (size_t)M * Kp * sizeof(bf16);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  if (smem > 48 * 1024) cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n_out = pair ? N / 2 : N;
  gemm_skinny_kernel<<<ceil_div(n_out, SK_WARPS), SK_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const bf16*>(A), lda, reinterpret_cast<const bf16*>(W), ldw, M, N, K, bias, residual, ld_res,
      out_f32, reinterpret_cast<bf16*>(out_bf16), ld_out, act);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
