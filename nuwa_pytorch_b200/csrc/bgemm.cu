// Batched small-matrix GEMM on the legacy tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32).
//
// Used by the attention BACKWARD of the dense Attention blocks (nuwa_pytorch.py:315-379 differentiated): the per-(batch,
// head) products  S = Q K^T,  dP' = dO V^T,  dQ = dS K,  dK = dS^T Q,  dV = P'^T dO  are 64..2560 x 64..264 matrices, far
// too small for the persistent tcgen05 kernel (gemm_tcgen05.cu) but numerous (B*H per layer) -- one launch covers them all
// (grid.z = batch1 * batch2, independent strides for both batch levels so heads can live inside a row).
//
//   C[m][n] = alpha * sum_k A(m,k) * B(k,n)   ( + C when accumulate )
//   A(m,k) = a_trans ? A[k*lda + m] : A[m*lda + k]         (the contiguous index must be padded to a multiple of 8 elements
//   B(k,n) = b_trans ? B[k*ldb + n] : B[n*ldb + k]          and 16-byte aligned: tiles are staged with 16-byte loads)
//
// CTA tile 64 x 64 x 32, 4 warps (2 x 2), each warp 32 x 32 = 2 x 4 mma tiles; operands staged in shared memory in their
// native orientation and read with ldmatrix / ldmatrix.trans, so no operand is ever transposed in HBM.
#include "common.cuh"
#include "kernels.h"

namespace nuwa {

namespace {

constexpr int BG_M = 64, BG_N = 64, BG_K = 32;
constexpr int PITCH_K = BG_K + 8;   // [64][32] tiles (k contiguous): 80-byte rows, conflict-free ldmatrix
constexpr int PITCH_MN = BG_M + 8;  // [32][64] tiles (m/n contiguous): 144-byte rows

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = smem_u32(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = smem_u32(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-byte chunk (8 bf16) of a row-major matrix with zero fill outside [0,rows) x [0,cols)
__device__ __forceinline__ uint4 load_chunk(const bf16* base, long long ld, int r, int c, int rows, int cols) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (r < rows && c < cols) {
    if (c + 8 <= cols) {
      v = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ld + c));
    } else {  // ragged tail (only when the caller's pad columns do not exist): element-wise
      unsigned short e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        e[i] = (c + i < cols) ? __ldg(reinterpret_cast<const unsigned short*>(base + (long long)r * ld + c + i)) : 0;
      v.x = e[0] | ((uint32_t)e[1] << 16); v.y = e[2] | ((uint32_t)e[3] << 16);
      v.z = e[4] | ((uint32_t)e[5] << 16); v.w = e[6] | ((uint32_t)e[7] << 16);
    }
  }
  return v;
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(128) bgemm_kernel(const nuwa_bgemm_params p) {
  __shared__ __align__(16) bf16 As[AT ? BG_K * PITCH_MN : BG_M * PITCH_K];
  __shared__ __align__(16) bf16 Bs[BT ? BG_K * PITCH_MN : BG_N * PITCH_K];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * BG_M, n0 = blockIdx.x * BG_N;
  const int i1 = blockIdx.z / p.batch2, i2 = blockIdx.z - i1 * p.batch2;
  const bf16* A = reinterpret_cast<const bf16*>(p.A) + i1 * p.a_s1 + i2 * p.a_s2;
  const bf16* B = reinterpret_cast<const bf16*>(p.B) + i1 * p.b_s1 + i2 * p.b_s2;

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  uint4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int ch = tid + s * 128;  // 256 chunks per tile
      if (AT) {  // A stored [k][m]: tile rows = k (32), 8 chunks of m per row
        const int kr = ch >> 3, mc = (ch & 7) * 8;
        ra[s] = load_chunk(A, p.lda, k0 + kr, m0 + mc, p.K, p.M);
      } else {   // A stored [m][k]: tile rows = m (64), 4 chunks of k per row
        const int mr = ch >> 2, kc = (ch & 3) * 8;
        ra[s] = load_chunk(A, p.lda, m0 + mr, k0 + kc, p.M, p.K);
      }
      if (BT) {  // B stored [k][n]
        const int kr = ch >> 3, nc = (ch & 7) * 8;
        rb[s] = load_chunk(B, p.ldb, k0 + kr, n0 + nc, p.K, p.N);
      } else {   // B stored [n][k]
        const int nr = ch >> 2, kc = (ch & 3) * 8;
        rb[s] = load_chunk(B, p.ldb, n0 + nr, k0 + kc, p.N, p.K);
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int ch = tid + s * 128;
      if (AT) *reinterpret_cast<uint4*>(&As[(ch >> 3) * PITCH_MN + (ch & 7) * 8]) = ra[s];
      else *reinterpret_cast<uint4*>(&As[(ch >> 2) * PITCH_K + (ch & 3) * 8]) = ra[s];
      if (BT) *reinterpret_cast<uint4*>(&Bs[(ch >> 3) * PITCH_MN + (ch & 7) * 8]) = rb[s];
      else *reinterpret_cast<uint4*>(&Bs[(ch >> 2) * PITCH_K + (ch & 3) * 8]) = rb[s];
    }
  };

  const int ktiles = (p.K + BG_K - 1) / BG_K;
  gload(0);
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();  // previous tile fully consumed
    sstore();
    __syncthreads();
    if (kt + 1 < ktiles) gload((kt + 1) * BG_K);  // global loads of the next tile overlap the MMAs below
#pragma unroll
    for (int ks = 0; ks < BG_K / 16; ++ks) {
      uint32_t af[2][4], bfr[2][4];
      const int id = lane >> 3, l8 = lane & 7;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mb = wm * 32 + mi * 16;
        if (AT) ldsm_x4_t(af[mi], &As[(ks * 16 + l8 + 8 * (id >> 1)) * PITCH_MN + mb + 8 * (id & 1)]);
        else ldsm_x4(af[mi], &As[(mb + l8 + 8 * (id & 1)) * PITCH_K + ks * 16 + 8 * (id >> 1)]);
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {  // two n8 tiles per ldmatrix.x4
        const int nb = wn * 32 + nj * 16;
        if (BT) ldsm_x4_t(bfr[nj], &Bs[(ks * 16 + l8 + 8 * (id & 1)) * PITCH_MN + nb + 8 * (id >> 1)]);
        else ldsm_x4(bfr[nj], &Bs[(nb + l8 + 8 * (id >> 1)) * PITCH_K + ks * 16 + 8 * (id & 1)]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nj = 0; nj < 2; ++nj) {
          mma_bf16(acc[mi][nj * 2 + 0], af[mi], bfr[nj][0], bfr[nj][1]);
          mma_bf16(acc[mi][nj * 2 + 1], af[mi], bfr[nj][2], bfr[nj][3]);
        }
    }
  }

  // ---- epilogue: c0,c1 = (row g, cols 2t,2t+1) ; c2,c3 = (row g+8, ...) ----
  const int g = lane >> 2, t = lane & 3;
  float* Cf = p.c_bf16 ? nullptr : reinterpret_cast<float*>(p.C) + i1 * p.c_s1 + i2 * p.c_s2;
  bf16* Cb = p.c_bf16 ? reinterpret_cast<bf16*>(p.C) + i1 * p.c_s1 + i2 * p.c_s2 : nullptr;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = m0 + wm * 32 + mi * 16 + g + hh * 8;
        const int n = n0 + wn * 32 + nt * 8 + 2 * t;
        if (m >= p.M || n >= p.N) continue;
        float v0 = acc[mi][nt][hh * 2] * p.alpha, v1 = acc[mi][nt][hh * 2 + 1] * p.alpha;
        const long long off = (long long)m * p.ldc + n;
        if (Cb != nullptr) {
          if (n + 1 < p.N && ((off & 1) == 0)) *reinterpret_cast<uint32_t*>(Cb + off) = pack_bf16x2(v0, v1);
          else {
            Cb[off] = __float2bfloat16(v0);
            if (n + 1 < p.N) Cb[off + 1] = __float2bfloat16(v1);
          }
        } else {
          if (p.accumulate) {
            v0 += Cf[off];
            if (n + 1 < p.N) v1 += Cf[off + 1];
          }
          Cf[off] = v0;
          if (n + 1 < p.N) Cf[off + 1] = v1;
        }
      }
}

}  // namespace

int bgemm(const nuwa_bgemm_params& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch1 <= 0 || p.batch2 <= 0) return NUWA_ERR_INVALID;
  if ((p.lda % 8) || (p.ldb % 8)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) return NUWA_ERR_INVALID;
  if ((p.a_s1 % 8) || (p.a_s2 % 8) || (p.b_s1 % 8) || (p.b_s2 % 8)) return NUWA_ERR_INVALID;
  const long long nz = (long long)p.batch1 * p.batch2;
  if (nz > 65535) return NUWA_ERR_INVALID;
  dim3 grid(ceil_div(p.N, BG_N), ceil_div(p.M, BG_M), (unsigned)nz);
  if (p.a_trans) {
    if (p.b_trans) bgemm_kernel<true, true><<<grid, 128, 0, stream>>>(p);
    else bgemm_kernel<true, false><<<grid, 128, 0, stream>>>(p);
  } else {
    if (p.b_trans) bgemm_kernel<false, true><<<grid, 128, 0, stream>>>(p);
    else bgemm_kernel<false, false><<<grid, 128, 0, stream>>>(p);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
