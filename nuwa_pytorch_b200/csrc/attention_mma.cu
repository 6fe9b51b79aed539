// Dense attention core on tensor cores for short key sets (nk <= 256): text cross-attention of the
// video decoder (nuwa_pytorch.py:339-378), dense self-attention of the text encoder, VQGanAttention core
// (vqgan_vae.py:275-282).
//
// One CTA = 16 query tokens x ALL heads of one sample, warp w = head w:
//   phase 1  S_w = (Q_w K_w^T) * scale (+bias), mask, fp32 softmax in the mma accumulator registers;
//            the learned null key is an exact fp32 side column; P_w -> shared memory as bf16
//   phase 2  talking heads: P'[g] = sum_h W[g][h] P[h] mixes the heads in shared memory
//   phase 3  O_w = P'_w V_w (+ P'_null * null_v), bf16 out
//
// Operand fragments come straight from global/L2 (K/V of a sample are shared by all its query tiles and stay
// L2 resident).  A contraction does not care about the order of its index, so the contraction index is
// PERMUTED such that everything one lane feeds into the 4 k-steps of an m16n8k16 group is 32 contiguous
// bytes: lane t of a quad owns d in [DH/4*t, DH/4*(t+1)) for QK^T and keys [64*blk+16t, +16) for PV
// (k-step s, fragment slots {2t,2t+1,2t+8,2t+9}  <->  owned index 4s+{0,1,2,3}).  Loads are therefore
// 16-byte vectors covering whole 128-byte rows per quad instead of scattered 4-byte words.
// V is consumed through a transposed copy vT[b][h][d][j] (kv_transpose_kernel).
//
// This is the legacy mma.sync tensor path on purpose: tiles are 16 x 8, far below a tcgen05 128-row atom, and
// the talking-heads mix forces all heads of a query tile to be resident at once (8 x 257 probabilities per
// query), which caps the query tile at 16-32 rows.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int MMA_MAXKB = 4;   // 64-key blocks -> nk <= 256
static constexpr int MMA_QT = 16;     // queries per CTA

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 2*KS consecutive 32-bit words (= 4*KS bf16) from a 16-byte aligned address
template <int KS>
__device__ __forceinline__ void load_words(const bf16* p, uint32_t (&w)[2 * KS]) {
  if constexpr (KS == 4) {
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    w[0] = u0.x; w[1] = u0.y; w[2] = u0.z; w[3] = u0.w;
    w[4] = u1.x; w[5] = u1.y; w[6] = u1.z; w[7] = u1.w;
  } else {
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = u0.x; w[1] = u0.y; w[2] = u0.z; w[3] = u0.w;
  }
}

// vT[b][h][d][jp] = v[b][j][h*dh + d]  (zero padded to jp keys)
__global__ void __launch_bounds__(256)
kv_transpose_kernel(const bf16* __restrict__ v, long long v_bs, int v_rs, bf16* __restrict__ vT, int B, int H, int dh,
                    int nk, int jp) {
  const long long total = (long long)B * H * dh * jp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % jp);
    const int d = (int)((i / jp) % dh);
    const int h = (int)((i / ((long long)jp * dh)) % H);
    const long long b = i / ((long long)jp * dh * H);
    vT[i] = j < nk ? v[b * v_bs + (long long)j * v_rs + h * dh + d] : __float2bfloat16(0.f);
  }
}

template <int DH>
__global__ void __launch_bounds__(256, 1) attn_dense_mma_kernel(const AttnParams p, const bf16* __restrict__ vT, int jp) {
  constexpr int KS = DH / 16;   // k-steps of the QK^T contraction
  constexpr int ND = DH / 8;    // output n-tiles of PV
  constexpr int SPAN = DH / 4;  // d values owned by one lane of a quad
  extern __shared__ __align__(16) uint8_t smem_mma[];
  const int H = p.H;
  const int PSTR = jp + 4;  // bf16 row pitch of P: 64-bit fragment reads are bank-conflict free
  bf16* P = reinterpret_cast<bf16*>(smem_mma);                                        // [H][16][PSTR]
  float* Pnull = reinterpret_cast<float*>(smem_mma + (size_t)H * MMA_QT * PSTR * 2);  // [H][16]
  float* Wt = Pnull + H * MMA_QT;                                                     // [H][H]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_per_b = (p.nq + MMA_QT - 1) / MMA_QT;
  const int b = blockIdx.x / tiles_per_b;
  const int q0 = (blockIdx.x - b * tiles_per_b) * MMA_QT;
  const int nk = p.nk_dense;
  const int nkb = jp / 64;  // 64-key blocks
  const bool has_null = p.null_k != nullptr;
  if (p.talk != nullptr)
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) Wt[i] = p.talk[i];

  const bf16* qb = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs + w * DH + SPAN * t;
  const bf16* kb = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs + w * DH + SPAN * t;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  const bool ok0 = r0 < p.nq, ok1 = r1 < p.nq;

  // ---------------- phase 1: scores + softmax (warp w = head w) ----------------
  uint32_t qa0[2 * KS], qa1[2 * KS];  // rows r0 / r1, this lane's d span
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) qa0[i] = qa1[i] = 0u;
  if (ok0) load_words<KS>(qb + (long long)r0 * p.q_rs, qa0);
  if (ok1) load_words<KS>(qb + (long long)r1 * p.q_rs, qa1);

  float s[8 * MMA_MAXKB][4];
#pragma unroll
  for (int nt = 0; nt < 8 * MMA_MAXKB; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    if (nt < 8 * nkb) {
      const int j = nt * 8 + g;  // key row this lane feeds into the B fragment
      uint32_t kw[2 * KS];
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) kw[i] = 0u;
      if (j < nk) load_words<KS>(kb + (long long)j * p.k_rs, kw);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        mma_bf16_16816(s[nt], qa0[2 * ks], qa1[2 * ks], qa0[2 * ks + 1], qa1[2 * ks + 1], kw[2 * ks], kw[2 * ks + 1]);
    }
  }
  const float hs = p.qscale * (p.head_scale != nullptr ? p.head_scale[w] : 1.0f);
  const unsigned char* km = p.key_mask != nullptr ? p.key_mask + (long long)b * p.mask_bs : nullptr;
  float m0 = -FLT_MAX, m1 = -FLT_MAX;
#pragma unroll
  for (int nt = 0; nt < 8 * MMA_MAXKB; ++nt) {
    if (nt < 8 * nkb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = nt * 8 + 2 * t + (e & 1);
        const int row = (e < 2) ? r0 : r1;
        float v = s[nt][e] * hs;
        if (p.bias != nullptr && j < nk && row < p.nq)
          v += p.bias[((long long)w * p.bias_nq + (p.t0 + row)) * p.bias_nk + j];
        if (j >= nk || (km != nullptr && km[j] == 0)) v = -FLT_MAX;
        s[nt][e] = v;
      }
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
  }
  // exact fp32 null-key logit (always visible): partial dot over this lane's d span, quad reduce
  float sn0 = 0.f, sn1 = 0.f;
  if (has_null) {
    const float* nkp = p.null_k + w * DH + SPAN * t;
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) {
      const float2 a0 = unpack_bf16x2(qa0[i]), a1 = unpack_bf16x2(qa1[i]);
      sn0 += a0.x * nkp[2 * i] + a0.y * nkp[2 * i + 1];
      sn1 += a1.x * nkp[2 * i] + a1.y * nkp[2 * i + 1];
    }
    sn0 += __shfl_xor_sync(0xffffffffu, sn0, 1); sn0 += __shfl_xor_sync(0xffffffffu, sn0, 2);
    sn1 += __shfl_xor_sync(0xffffffffu, sn1, 1); sn1 += __shfl_xor_sync(0xffffffffu, sn1, 2);
    sn0 *= hs; sn1 *= hs;
    m0 = fmaxf(m0, sn0);
    m1 = fmaxf(m1, sn1);
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8 * MMA_MAXKB; ++nt) {
    if (nt < 8 * nkb) {
      s[nt][0] = __expf(s[nt][0] - m0); s[nt][1] = __expf(s[nt][1] - m0);
      s[nt][2] = __expf(s[nt][2] - m1); s[nt][3] = __expf(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float en0 = 0.f, en1 = 0.f;
  if (has_null) {
    en0 = __expf(sn0 - m0);
    en1 = __expf(sn1 - m1);
    l0 += en0;
    l1 += en1;
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  bf16* Pw = P + (size_t)w * MMA_QT * PSTR;
#pragma unroll
  for (int nt = 0; nt < 8 * MMA_MAXKB; ++nt) {
    if (nt < 8 * nkb) {
      *reinterpret_cast<uint32_t*>(Pw + (size_t)g * PSTR + nt * 8 + 2 * t) = pack_bf16x2(s[nt][0] * i0, s[nt][1] * i0);
      *reinterpret_cast<uint32_t*>(Pw + (size_t)(g + 8) * PSTR + nt * 8 + 2 * t) = pack_bf16x2(s[nt][2] * i1, s[nt][3] * i1);
    }
  }
  if (t == 0) {
    Pnull[w * MMA_QT + g] = en0 * i0;
    Pnull[w * MMA_QT + g + 8] = en1 * i1;
  }
  __syncthreads();

  // ---------------- phase 2: talking heads (mix the H probability rows of every (query, key)) ----------------
  if (p.talk != nullptr) {
    const int half = jp / 2;
    for (int i = threadIdx.x; i < MMA_QT * half; i += blockDim.x) {
      const int q = i / half, jj = (i - q * half) * 2;
      float2 pin[8];
      for (int h = 0; h < H; ++h)
        pin[h] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P + ((size_t)h * MMA_QT + q) * PSTR + jj));
      for (int gh = 0; gh < H; ++gh) {
        float ax = 0.f, ay = 0.f;
        for (int h = 0; h < H; ++h) {
          ax = fmaf(Wt[gh * H + h], pin[h].x, ax);
          ay = fmaf(Wt[gh * H + h], pin[h].y, ay);
        }
        *reinterpret_cast<uint32_t*>(P + ((size_t)gh * MMA_QT + q) * PSTR + jj) = pack_bf16x2(ax, ay);
      }
    }
    if (threadIdx.x < MMA_QT) {
      const int q = threadIdx.x;
      float pin[8];
      for (int h = 0; h < H; ++h) pin[h] = Pnull[h * MMA_QT + q];
      for (int gh = 0; gh < H; ++gh) {
        float a = 0.f;
        for (int h = 0; h < H; ++h) a = fmaf(Wt[gh * H + h], pin[h], a);
        Pnull[gh * MMA_QT + q] = a;
      }
    }
    __syncthreads();
  }

  // ---------------- phase 3: O_w = P'_w V_w  (keys contracted in lane-owned 16-key spans) ----------------
  float o[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  const bf16* vtb = vT + ((long long)b * H + w) * DH * jp + 16 * t;
  for (int blk = 0; blk < nkb; ++blk) {
    // A fragments of the 4 k-steps of this 64-key block: P'[row][64 blk + 16 t + 4 s + {0,1 | 2,3}]
    uint2 pa0[4], pa1[4];
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
      pa0[sidx] = *reinterpret_cast<const uint2*>(Pw + (size_t)g * PSTR + blk * 64 + 16 * t + 4 * sidx);
      pa1[sidx] = *reinterpret_cast<const uint2*>(Pw + (size_t)(g + 8) * PSTR + blk * 64 + 16 * t + 4 * sidx);
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      uint32_t vw[8];
      load_words<4>(vtb + (long long)(nd * 8 + g) * jp + blk * 64, vw);
#pragma unroll
      for (int sidx = 0; sidx < 4; ++sidx)
        mma_bf16_16816(o[nd], pa0[sidx].x, pa1[sidx].x, pa0[sidx].y, pa1[sidx].y, vw[2 * sidx], vw[2 * sidx + 1]);
    }
  }
  bf16* ob = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs + w * DH;
  const float pn0 = has_null ? Pnull[w * MMA_QT + g] : 0.f, pn1 = has_null ? Pnull[w * MMA_QT + g + 8] : 0.f;
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int d = nd * 8 + 2 * t;
    float v0 = o[nd][0], v1 = o[nd][1], v2 = o[nd][2], v3 = o[nd][3];
    if (has_null) {
      const float nv0 = p.null_v[w * DH + d], nv1 = p.null_v[w * DH + d + 1];
      v0 = fmaf(pn0, nv0, v0); v1 = fmaf(pn0, nv1, v1);
      v2 = fmaf(pn1, nv0, v2); v3 = fmaf(pn1, nv1, v3);
    }
    if (ok0) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * p.o_rs + d) = pack_bf16x2(v0, v1);
    if (ok1) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * p.o_rs + d) = pack_bf16x2(v2, v3);
  }
}

// Returns NUWA_ERR_INVALID when the shape is outside this kernel's envelope (caller falls back to the
// generic CUDA-core attention kernel).  vT_ws: workspace of B*H*dh*roundup(nk,64) bf16.
int attn_dense_mma(const AttnParams& p, int nk, void* vT_ws, cudaStream_t stream) {
  if (p.H > 8 || (p.dh != 64 && p.dh != 32) || nk <= 0 || nk > 64 * MMA_MAXKB || vT_ws == nullptr) return NUWA_ERR_INVALID;
  // 16-byte vector loads of q / k rows
  if ((p.q_rs % 8) || (p.k_rs % 8) || (p.q_bs % 8) || (p.k_bs % 8) || (p.o_rs & 1)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.k) & 15)) return NUWA_ERR_INVALID;
  const int jp = (nk + 63) / 64 * 64;
  bf16* vT = reinterpret_cast<bf16*>(vT_ws);
  const long long total = (long long)p.B * p.H * p.dh * jp;
  int tgrid = (int)((total + 255) / 256);
  if (tgrid > 148 * 16) tgrid = 148 * 16;
  kv_transpose_kernel<<<tgrid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(p.v), p.v_bs, p.v_rs, vT, p.B, p.H, p.dh,
                                                  nk, jp);
  NUWA_CHECK_LAUNCH();
  AttnParams q = p;
  q.nk_dense = nk;
  const size_t smem = (size_t)p.H * MMA_QT * (jp + 4) * 2 + (size_t)p.H * MMA_QT * 4 + (size_t)p.H * p.H * 4;
  const int grid = p.B * ((p.nq + MMA_QT - 1) / MMA_QT);
  if (p.dh == 64) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_dense_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_dense_mma_kernel<64><<<grid, 32 * p.H, smem, stream>>>(q, vT, jp);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_dense_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_dense_mma_kernel<32><<<grid, 32 * p.H, smem, stream>>>(q, vT, jp);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
