// Skinny-M GEMM for the decode step of generate(): out[M,N] = act(A[M,K] W[N,K]^T + bias) + residual with
// M <= 32 (one row per sample).  Such a product streams every weight exactly once and does 2*M FLOPs per
// weight byte pair: it is HBM/L2-bandwidth and latency bound, a 128-row tensor-core tile would idle 94 % of
// its rows and occupy only N/64 CTAs.  Here every warp owns one output column (or one value/gate pair), the
// 32 lanes split K with 16-byte loads, A sits in shared memory, and N/8 CTAs cover the chip.
// The kernel is a chain of memory latencies, so ALL weight loads of a warp (up to SK_MAXC 16-byte chunks per
// lane) are issued before the activation tile is staged and before anything depends on them.
// Same operand conventions and fused epilogues as gemm_tcgen05_kernel (incl. pair-packed GLU/GEGLU rows).
#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int SK_WARPS = 8;
static constexpr int SK_MR = 8;    // rows accumulated per pass
static constexpr int SK_MAXC = 8;  // 16-byte chunks per lane held in registers -> K <= 2048 on the fast path

__device__ __forceinline__ float dot8(const uint4& a, const uint4& w) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  return a0.x * w0.x + a0.y * w0.y + a1.x * w1.x + a1.y * w1.y + a2.x * w2.x + a2.y * w2.y + a3.x * w3.x + a3.y * w3.y;
}

template <bool PAIR>
__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, int M, int N, int K,
                   const float* __restrict__ bias, const float* __restrict__ residual, int ld_res,
                   float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int ld_out, int act) {
  extern __shared__ __align__(16) uint8_t smem_sk[];
  bf16* As = reinterpret_cast<bf16*>(smem_sk);  // [M][K]   (K % 8 == 0)
  const int k8 = K / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = PAIR ? N / 2 : N;
  const int col = blockIdx.x * SK_WARPS + warp;  // output column of this warp
  const bool active = col < n_out;
  const int wrow = PAIR ? (col / 16) * 32 + (col % 16) : col;  // weight row(s) feeding this column
  const bf16* w0 = W + (long long)(active ? wrow : 0) * ldw;
  const bf16* w1 = w0 + (long long)16 * ldw;
  // ---- 1. all weight chunks of this lane in flight ----
  uint4 wv[SK_MAXC], wg[PAIR ? SK_MAXC : 1];
#pragma unroll
  for (int i = 0; i < SK_MAXC; ++i) {
    const int c = (lane + 32 * i) * 8;
    wv[i] = make_uint4(0u, 0u, 0u, 0u);
    if (PAIR) wg[i] = make_uint4(0u, 0u, 0u, 0u);
    if (active && c < K) {
      wv[i] = __ldg(reinterpret_cast<const uint4*>(w0 + c));
      if (PAIR) wg[i] = __ldg(reinterpret_cast<const uint4*>(w1 + c));
    }
  }
  // ---- 2. stage the activation rows ----
  for (int i = threadIdx.x; i < M * k8; i += blockDim.x) {
    const int r = i / k8, c = (i - r * k8) * 8;
    *reinterpret_cast<uint4*>(As + (long long)r * K + c) = *reinterpret_cast<const uint4*>(A + (long long)r * lda + c);
  }
  __syncthreads();
  if (!active) return;
  // ---- 3. dot products ----
  for (int m0 = 0; m0 < M; m0 += SK_MR) {
    float acc[SK_MR], accg[PAIR ? SK_MR : 1];
#pragma unroll
    for (int r = 0; r < SK_MR; ++r) {
      acc[r] = 0.f;
      if (PAIR) accg[r] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < SK_MAXC; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < K) {
#pragma unroll
        for (int r = 0; r < SK_MR; ++r)
          if (m0 + r < M) {
            const uint4 av = *reinterpret_cast<const uint4*>(As + (long long)(m0 + r) * K + c);
            acc[r] += dot8(av, wv[i]);
            if (PAIR) accg[r] += dot8(av, wg[i]);
          }
      }
    }
    // K beyond the register-resident chunks (rare): stream the rest
    for (int c = (lane + 32 * SK_MAXC) * 8; c < K; c += 256) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(w0 + c));
      uint4 xg = make_uint4(0u, 0u, 0u, 0u);
      if (PAIR) xg = __ldg(reinterpret_cast<const uint4*>(w1 + c));
#pragma unroll
      for (int r = 0; r < SK_MR; ++r)
        if (m0 + r < M) {
          const uint4 av = *reinterpret_cast<const uint4*>(As + (long long)(m0 + r) * K + c);
          acc[r] += dot8(av, x);
          if (PAIR) accg[r] += dot8(av, xg);
        }
    }
    // ---- 4. reduce over the lanes; lane r finishes row r ----
    float mine = 0.f, mineg = 0.f;
#pragma unroll
    for (int r = 0; r < SK_MR; ++r) {
      const float v = warp_sum(acc[r]);
      const float g = PAIR ? warp_sum(accg[r]) : 0.f;
      if (lane == r) { mine = v; mineg = g; }
    }
    if (lane < SK_MR && m0 + lane < M) {
      const long long m = m0 + lane;
      float v = mine;
      if (PAIR) {
        float g = mineg;
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

int gemm_skinny(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act,
                cudaStream_t stream) {
  if (M <= 0 || M > 32 || N <= 0 || K <= 0) return NUWA_ERR_INVALID;
  if ((lda % 8) || (ldw % 8) || (K % 8)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return NUWA_ERR_INVALID;
  const bool pair = (act == ACT_GLU || act == ACT_GEGLU);
  if (pair && (N % 32)) return NUWA_ERR_INVALID;
  const size_t smem = (size_t)M * K * sizeof(bf16);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  const int n_out = pair ? N / 2 : N;
  const int grid = ceil_div(n_out, SK_WARPS);
  const bf16* a = reinterpret_cast<const bf16*>(A);
  const bf16* w = reinterpret_cast<const bf16*>(W);
  bf16* ob = reinterpret_cast<bf16*>(out_bf16);
  if (pair) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(gemm_skinny_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_skinny_kernel<true><<<grid, SK_WARPS * 32, smem, stream>>>(a, lda, w, ldw, M, N, K, bias, residual, ld_res,
                                                                    out_f32, ob, ld_out, act);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(gemm_skinny_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_skinny_kernel<false><<<grid, SK_WARPS * 32, smem, stream>>>(a, lda, w, ldw, M, N, K, bias, residual, ld_res,
                                                                     out_f32, ob, ld_out, act);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
