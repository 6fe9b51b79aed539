// Dense attention for ONE query per sample, forward and backward: the bos query of SparseCross2DNA (nuwa_pytorch.py:828-844:
// [null key | every context token], context mask, NO talking heads) in full teacher-forced passes.  A single query has no
// query parallelism, so the generic paths ran it as a chain of tiny launches (decode kernel 102 us forward; K/V repack, five
// batched GEMMs, the wide row kernel and a split in the backward: ~0.25 ms per layer for 4 x 8 rows of 769 logits -- 12 % of
// the NUWASketch step).  Here: one CTA per (sample, head), thread = key for the dot products, block reductions for the
// softmax statistics, channel-parallel reductions for the outputs; the backward writes its dk / dv rows as fp32 (they are
// the base the key-centric pass of the windowed queries adds to) and adds the null key / value gradients with atomics.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {
namespace {

constexpr int Q1_THREADS = 256;
constexpr int Q1_DH = 64;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < Q1_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__device__ __forceinline__ float dot64(const float* q, const bf16* row) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < Q1_DH / 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(row) + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    const float* w = q + i * 8;
    s = fmaf(a.x, w[0], s); s = fmaf(a.y, w[1], s); s = fmaf(b.x, w[2], s); s = fmaf(b.y, w[3], s);
    s = fmaf(c.x, w[4], s); s = fmaf(c.y, w[5], s); s = fmaf(d.x, w[6], s); s = fmaf(d.y, w[7], s);
  }
  return s;
}

struct Q1Args {
  const bf16* q;    // [B] rows (q_bs), H * 64 channels
  const bf16* k;    // [B][nk] rows (k_bs / k_rs)
  const bf16* v;
  const bf16* dO;   // backward: [B] rows (do_bs)
  long long q_bs, k_bs, v_bs, do_bs, o_bs;
  int k_rs, v_rs;
  int B, H, nk;
  float qscale;
  const float* null_k;   // fp32 [H * 64] or NULL
  const float* null_v;
  const unsigned char* key_mask;
  int mask_bs;
  bf16* o;          // forward: [B] rows (o_bs)
  bf16* dq;         // backward: [B] rows (o_bs)
  float* dk;        // backward: fp32 [B][nk] rows (dkv_bs / dkv_rs)
  float* dv;
  long long dkv_bs;
  int dkv_rs;
  float* dnull_k;   // fp32 [H * 64], added to
  float* dnull_v;
};

// probabilities of (sample b, head h) into P[0 .. nk] (slot 0 = null key, P[0] = 0 without one); returns nothing else
__device__ __forceinline__ void q1_probs(const Q1Args& p, int b, int h, const float* qs, float* P, float* red) {
  const bool has_null = p.null_k != nullptr;
  const bf16* kb = p.k + (long long)b * p.k_bs + h * Q1_DH;
  float m = -FLT_MAX;
  for (int j = threadIdx.x; j <= p.nk; j += Q1_THREADS) {
    float s;
    if (j == 0) {
      s = -FLT_MAX;
      if (has_null) {
        s = 0.f;
        for (int c = 0; c < Q1_DH; ++c) s = fmaf(qs[c], __ldg(p.null_k + h * Q1_DH + c), s);
      }
    } else {
      const bool live = p.key_mask == nullptr || p.key_mask[(long long)b * p.mask_bs + j - 1] != 0;
      s = live ? dot64(qs, kb + (long long)(j - 1) * p.k_rs) : -FLT_MAX;
    }
    P[j] = s;
    m = fmaxf(m, s);
  }
  m = block_reduce(m, red, true);
  float l = 0.f;
  for (int j = threadIdx.x; j <= p.nk; j += Q1_THREADS) {
    const float e = P[j] == -FLT_MAX ? 0.f : __expf(P[j] - m);
    P[j] = e;
    l += e;
  }
  l = block_reduce(l, red, false);
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  for (int j = threadIdx.x; j <= p.nk; j += Q1_THREADS) P[j] *= inv;
  __syncthreads();
}

// out[c] = w[0] * nullrow[c] + sum_j w[j] * rows[j - 1][c] for the 64 channels of head h; thread = (key group, channel)
__device__ __forceinline__ void q1_weighted_rows(const float* w, const bf16* rows, int rs, const float* nullrow, int nk,
                                                 float* part, float* out) {
  const int c = threadIdx.x & (Q1_DH - 1), kg = threadIdx.x / Q1_DH;   // 4 key groups
  float acc = 0.f;
  for (int j = 1 + kg; j <= nk; j += Q1_THREADS / Q1_DH) acc = fmaf(w[j], __bfloat162float(rows[(long long)(j - 1) * rs + c]), acc);
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < Q1_DH) {
    float a = nullrow != nullptr ? w[0] * __ldg(nullrow + c) : 0.f;
#pragma unroll
    for (int g = 0; g < Q1_THREADS / Q1_DH; ++g) a += part[g * Q1_DH + c];
    out[c] = a;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(Q1_THREADS) attn_q1_fwd_kernel(const Q1Args p) {
  extern __shared__ float sm_q1[];
  float* P = sm_q1;                    // [nk + 1]
  float* qs = P + ((p.nk + 1 + 3) & ~3);
  float* part = qs + Q1_DH;            // [256]
  float* out = part + Q1_THREADS;      // [64]
  float* red = out + Q1_DH;            // [8]
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  if (threadIdx.x < Q1_DH) qs[threadIdx.x] = __bfloat162float(p.q[(long long)b * p.q_bs + h * Q1_DH + threadIdx.x]) * p.qscale;
  __syncthreads();
  q1_probs(p, b, h, qs, P, red);
  q1_weighted_rows(P, p.v + (long long)b * p.v_bs + h * Q1_DH, p.v_rs, p.null_v != nullptr ? p.null_v + h * Q1_DH : nullptr, p.nk,
                   part, out);
  if (threadIdx.x < Q1_DH) p.o[(long long)b * p.o_bs + h * Q1_DH + threadIdx.x] = __float2bfloat16(out[threadIdx.x]);
}

__global__ void __launch_bounds__(Q1_THREADS) attn_q1_bwd_kernel(const Q1Args p) {
  extern __shared__ float sm_q1[];
  float* P = sm_q1;                                  // [nk + 1] probabilities
  float* dS = P + ((p.nk + 1 + 3) & ~3);             // [nk + 1] dP, then dS
  float* qs = dS + ((p.nk + 1 + 3) & ~3);            // scaled q
  float* qraw = qs + Q1_DH;                          // unscaled q
  float* dOs = qraw + Q1_DH;
  float* part = dOs + Q1_DH;
  float* out = part + Q1_THREADS;
  float* red = out + Q1_DH;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const bool has_null = p.null_k != nullptr;
  if (threadIdx.x < Q1_DH) {
    const float qv = __bfloat162float(p.q[(long long)b * p.q_bs + h * Q1_DH + threadIdx.x]);
    qraw[threadIdx.x] = qv;
    qs[threadIdx.x] = qv * p.qscale;
    dOs[threadIdx.x] = __bfloat162float(p.dO[(long long)b * p.do_bs + h * Q1_DH + threadIdx.x]);
  }
  __syncthreads();
  q1_probs(p, b, h, qs, P, red);
  // dP_j = dO . v_j ; delta = sum_j P_j dP_j
  const bf16* vb = p.v + (long long)b * p.v_bs + h * Q1_DH;
  float dl = 0.f;
  for (int j = threadIdx.x; j <= p.nk; j += Q1_THREADS) {
    float d = 0.f;
    if (P[j] != 0.f) {
      if (j == 0) {
        for (int c = 0; c < Q1_DH; ++c) d = fmaf(dOs[c], __ldg(p.null_v + h * Q1_DH + c), d);
      } else {
        d = dot64(dOs, vb + (long long)(j - 1) * p.v_rs);
      }
    }
    dS[j] = d;
    dl = fmaf(P[j], d, dl);
  }
  dl = block_reduce(dl, red, false);
  for (int j = threadIdx.x; j <= p.nk; j += Q1_THREADS) dS[j] = P[j] * (dS[j] - dl) * p.qscale;   // dS carries the logit scale
  __syncthreads();
  // dq = sum_j dS_j k_j (+ null key)
  q1_weighted_rows(dS, p.k + (long long)b * p.k_bs + h * Q1_DH, p.k_rs, has_null ? p.null_k + h * Q1_DH : nullptr, p.nk, part, out);
  if (threadIdx.x < Q1_DH) p.dq[(long long)b * p.o_bs + h * Q1_DH + threadIdx.x] = __float2bfloat16(out[threadIdx.x]);
  // dk_j = dS_j q ; dv_j = P_j dO   (fp32 rows; masked keys get exact zeros); thread = (key group, channel)
  {
    const int c = threadIdx.x & (Q1_DH - 1), kg = threadIdx.x / Q1_DH;
    float* dkb = p.dk + (long long)b * p.dkv_bs + h * Q1_DH + c;
    float* dvb = p.dv + (long long)b * p.dkv_bs + h * Q1_DH + c;
    const float qc = qraw[c], dc = dOs[c];
    for (int j = 1 + kg; j <= p.nk; j += Q1_THREADS / Q1_DH) {
      dkb[(long long)(j - 1) * p.dkv_rs] = dS[j] * qc;
      dvb[(long long)(j - 1) * p.dkv_rs] = P[j] * dc;
    }
    if (has_null && kg == 0) {
      if (p.dnull_k != nullptr) atomicAdd(p.dnull_k + h * Q1_DH + c, dS[0] * qc);
      if (p.dnull_v != nullptr) atomicAdd(p.dnull_v + h * Q1_DH + c, P[0] * dc);
    }
  }
}

bool q1_args_ok(const Q1Args& a) {
  if (a.B <= 0 || a.H <= 0 || a.nk <= 0 || a.nk > 8192) return false;
  if ((a.null_k == nullptr) != (a.null_v == nullptr)) return false;
  if ((a.k_rs % 8) || (a.v_rs % 8) || (a.k_bs % 8) || (a.v_bs % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(a.k) & 15) || (reinterpret_cast<uintptr_t>(a.v) & 15)) return false;
  return true;
}

}  // namespace

// Forward: p->q / p->o point at the one query / output row of every sample (strides q_bs / o_bs), p->nq == 1, no talking
// heads / bias / per-head scale, dh == 64.  NUWA_ERR_INVALID outside that envelope (nothing launched).
int attn_dense_q1(const AttnParams& p, int nk, cudaStream_t stream) {
  if (p.nq != 1 || p.dh != Q1_DH || p.talk != nullptr || p.bias != nullptr || p.head_scale != nullptr || p.t0_ptr != nullptr)
    return NUWA_ERR_INVALID;
  Q1Args a = {};
  a.q = reinterpret_cast<const bf16*>(p.q); a.k = reinterpret_cast<const bf16*>(p.k); a.v = reinterpret_cast<const bf16*>(p.v);
  a.q_bs = p.q_bs; a.k_bs = p.k_bs; a.v_bs = p.v_bs; a.o_bs = p.o_bs; a.k_rs = p.k_rs; a.v_rs = p.v_rs;
  a.B = p.B; a.H = p.H; a.nk = nk; a.qscale = p.qscale;
  a.null_k = p.null_k; a.null_v = p.null_v; a.key_mask = p.key_mask; a.mask_bs = p.mask_bs;
  a.o = reinterpret_cast<bf16*>(p.o);
  if (!q1_args_ok(a) || a.o == nullptr) return NUWA_ERR_INVALID;
  const size_t smem = (((size_t)nk + 1 + 3) & ~(size_t)3) * 4 + (Q1_DH + Q1_THREADS + Q1_DH + 8) * 4;
  if (smem > 48 * 1024) return NUWA_ERR_INVALID;
  attn_q1_fwd_kernel<<<p.B * p.H, Q1_THREADS, smem, stream>>>(a);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// Backward of the same: dO / dq one row per sample (do_bs / dq_bs); dk, dv fp32 [B][nk] rows (dkv_bs / dkv_rs), WRITTEN (every
// row, zeros for masked keys); dnull_k / dnull_v fp32 [H * 64], ADDED to.
int attn_dense_q1_bwd(const AttnParams& p, int nk, const void* dO, long long do_bs, void* dq, long long dq_bs, float* dk, float* dv,
                      long long dkv_bs, int dkv_rs, float* dnull_k, float* dnull_v, cudaStream_t stream) {
  if (p.nq != 1 || p.dh != Q1_DH || p.talk != nullptr || p.bias != nullptr || p.head_scale != nullptr || p.t0_ptr != nullptr)
    return NUWA_ERR_INVALID;
  if (dO == nullptr || dq == nullptr || dk == nullptr || dv == nullptr) return NUWA_ERR_INVALID;
  Q1Args a = {};
  a.q = reinterpret_cast<const bf16*>(p.q); a.k = reinterpret_cast<const bf16*>(p.k); a.v = reinterpret_cast<const bf16*>(p.v);
  a.dO = reinterpret_cast<const bf16*>(dO);
  a.q_bs = p.q_bs; a.k_bs = p.k_bs; a.v_bs = p.v_bs; a.do_bs = do_bs; a.o_bs = dq_bs; a.k_rs = p.k_rs; a.v_rs = p.v_rs;
  a.B = p.B; a.H = p.H; a.nk = nk; a.qscale = p.qscale;
  a.null_k = p.null_k; a.null_v = p.null_v; a.key_mask = p.key_mask; a.mask_bs = p.mask_bs;
  a.dq = reinterpret_cast<bf16*>(dq); a.dk = dk; a.dv = dv; a.dkv_bs = dkv_bs; a.dkv_rs = dkv_rs;
  a.dnull_k = dnull_k; a.dnull_v = dnull_v;
  if (!q1_args_ok(a)) return NUWA_ERR_INVALID;
  const size_t smem = 2 * (((size_t)nk + 1 + 3) & ~(size_t)3) * 4 + (3 * Q1_DH + Q1_THREADS + Q1_DH + 8) * 4;
  if (smem > 48 * 1024) return NUWA_ERR_INVALID;
  attn_q1_bwd_kernel<<<p.B * p.H, Q1_THREADS, smem, stream>>>(a);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
