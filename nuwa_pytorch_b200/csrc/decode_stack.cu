// Persistent decode-step kernel for generate(): ONE cooperative launch walks a whole decoder stack
// (every SandwichNorm-wrapped Sparse3DNA / cross-attention / FeedForward sub-block, the final
// StableLayerNorm and optionally the to_logits product) for the single new position of every sample.
//
// Why: a decode step of the 64-layer reversible decoder is 256 sub-blocks x ~4 kernels, each a few
// microseconds of work on M = batch rows; as separate launches (even replayed from a CUDA graph) the step
// is bound by launch + dependency latency (~7.6 us per node, 14.6 ms per token), 60x above the time the
// 806 MB of bf16 weights need to stream from HBM.  Here the chip stays resident: 148 CTAs x 16 warps,
// phases separated by a device-wide barrier (one L2 atomic + acquire spin), and
//   * the residual streams live in EVERY CTA's shared memory (each CTA redundantly applies the post-norm +
//     residual + pre-norm of the B rows, so no extra barrier and no HBM round trip for the streams),
//   * every skinny product (q|k|v, q, GEGLU, out, FF-out, logits) is spread as (column, K-slice) tasks over
//     all warps of the chip; the weights of a warp's first task are requested BEFORE it waits at the
//     barrier, so the HBM latency of the weight stream is hidden behind the barrier,
//   * the new token's q|k|v go straight into the KV cache row (no append kernel), the ShiftVideoTokens
//     gather reads the persistent pre-norm cache, cross attention is split over (sample, head) CTAs in two
//     phases (scores | softmax + talking-heads row + PV) so that no SM streams more than one head.
// Rounding points are those of the per-kernel path (bf16 GEMM operands / q,k,v,o / GEGLU output, fp32
// accumulation, norms and streams), so both paths agree to fp32 summation order.
//
// Replaces, for one token step: Transformer / ReversibleTransformer.forward (nuwa_pytorch.py:1168-1182,
// :1289-1295, reversible.py:61-68,132-142) over SandwichNorm (:112-128), ShiftVideoTokens (:200-253),
// Sparse3DNA (:459-613), Attention (:315-379), FeedForward/GEGLU (:255-286), StableLayerNorm (:88-95) and
// to_logits (:1819).
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

typedef nuwa_decode_sub DecSub;
typedef nuwa_decode_params DecParams;

static constexpr int DS_THREADS = 512;
static constexpr int DS_WARPS = DS_THREADS / 32;
static constexpr int DS_MR = 8;      // rows accumulated per pass of a skinny product
static constexpr int DS_MAXC = 2;    // prefetched 16-byte weight chunks per lane and weight row
static constexpr int DS_LNV = 8;     // float4 per lane in the row norms -> D <= 1024

enum { DK_NORMAL = 0, DK_MASKED = 1, DK_ZERO = 2, DK_NULL = 3 };

// ------------------------------------------------------------------------------------------------
// shared-memory layout (same arithmetic on host and device)
// ------------------------------------------------------------------------------------------------
struct DsLayout {
  int streams, act, red, S, pm, keys, qs, wt, part, outs, lnp, shs, desc, total;
};
__host__ __device__ inline int ds_align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline DsLayout ds_layout(int B, int D, int kmax, int H, int dh, int jmax) {
  DsLayout L;
  int o = 0;
  L.streams = o; o += ds_align16(2 * B * D * 4);          // fp32 [2][B][D]
  L.act = o;     o += ds_align16(B * kmax * 2);           // bf16 [B][kmax] operand rows of the current product
  L.red = o;     o += ds_align16(DS_WARPS * 2 * DS_MR * 4);  // K-slice partial sums
  L.S = o;       o += ds_align16(H * jmax * 4);           // scores / probabilities [H][J]
  L.pm = o;      o += ds_align16(jmax * 4);               // mixed probabilities of one head
  L.keys = o;    o += ds_align16(jmax * 4);               // key list: (kind << 28) | row
  L.qs = o;      o += ds_align16(H * dh * 4);             // scaled query, fp32
  L.wt = o;      o += ds_align16(H * H * 4);              // talking-heads matrix
  L.part = o;    o += ds_align16(DS_THREADS * 2 * 4);     // PV partial sums
  L.outs = o;    o += ds_align16(H * dh * 4);             // attention output, fp32
  L.lnp = o;     o += ds_align16(4 * D * 4);              // post_w, post_b, pre_w, pre_b of the coming norms (cp.async)
  L.shs = o;     o += ds_align16(B * (D / 2) * 2);        // shifted channel halves of the ShiftVideoTokens gather
  L.desc = o;    o += 3 * ds_align16((int)sizeof(nuwa_decode_sub));  // previous / current / next sub-block descriptor
  L.total = o;
  return L;
}

// ------------------------------------------------------------------------------------------------
// device-wide barrier: monotonically increasing arrival counter (reset by the last CTA to leave the kernel)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct GridBar {
  unsigned* count;
  unsigned nblocks;
  unsigned target;
};

__device__ __forceinline__ void grid_barrier(GridBar& g) {
  __syncthreads();
  if (threadIdx.x == 0) {
    g.target += g.nblocks;
    // release: everything this CTA wrote (ordered before by the CTA barrier) is visible to whoever acquires the count
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(g.count) : "memory");
    long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_u32(g.count) < g.target) {
      ++spins;
      if (spins == 4096u) t0 = clock64();
      // watchdog: a protocol bug becomes a launch error instead of a hung GPU (~2 s)
      if (spins > 4096u && (spins & 4095u) == 0u && (clock64() - t0) > 4000000000LL) __trap();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// skinny products:  out[b][col] = sum_k As[b][k] * W[row(col)][k]      (As bf16 in shared memory)
// task = (column, K-slice); the S slices of a column sit in consecutive warps of one CTA.
// ------------------------------------------------------------------------------------------------
struct WPre {
  uint4 v[DS_MAXC];
  uint4 g[DS_MAXC];
};

__device__ __forceinline__ float dot8f(const uint4& a, const uint4& w) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  return a0.x * w0.x + a0.y * w0.y + a1.x * w1.x + a1.y * w1.y + a2.x * w2.x + a2.y * w2.y + a3.x * w3.x + a3.y * w3.y;
}

__device__ __forceinline__ int gemv_slice(int K, int S) { return ((K + S - 1) / S + 7) & ~7; }

struct GemvOut {
  float* f32;        // fp32 output or NULL
  bf16* b16;         // bf16 output or NULL
  long long ld;      // row stride (elements)
  const float* bias; // per output column or NULL
};

template <bool PAIR>
__device__ __forceinline__ void gemv_prefetch(WPre& w, const bf16* __restrict__ W, int ldw, int N, int K, int S,
                                              int gwarp, int lane) {
#pragma unroll
  for (int i = 0; i < DS_MAXC; ++i) {
    w.v[i] = make_uint4(0u, 0u, 0u, 0u);
    w.g[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (gwarp >= N * S) return;
  const int col = gwarp / S, ks = gwarp - col * S;
  const int slice = gemv_slice(K, S);
  const int k0 = ks * slice, k1 = min(K, k0 + slice);
  const int row = PAIR ? (col / 16) * 32 + (col % 16) : col;
  const bf16* w0 = W + (long long)row * ldw;
#pragma unroll
  for (int i = 0; i < DS_MAXC; ++i) {
    const int c = k0 + (lane + 32 * i) * 8;
    if (c < k1) {
      w.v[i] = __ldg(reinterpret_cast<const uint4*>(w0 + c));
      if (PAIR) w.g[i] = __ldg(reinterpret_cast<const uint4*>(w0 + (long long)16 * ldw + c));
    }
  }
}

// All threads of the CTA must call this (it contains __syncthreads when S > 1).
template <bool PAIR>
__device__ __forceinline__ void gemv_run(const WPre& wp, const bf16* __restrict__ W, int ldw, int N, int K, int S,
                                         const bf16* As, int lda_s, int B, float* red, const GemvOut& o, int gwarp,
                                         int total_warps, int warp, int lane) {
  const int ntasks = N * S;
  const int rounds = (ntasks + total_warps - 1) / total_warps;
  const int slice = gemv_slice(K, S);
  for (int r = 0; r < rounds; ++r) {
    const int task = gwarp + r * total_warps;
    const bool active = task < ntasks;
    const int col = active ? task / S : 0, ks = active ? task - col * S : 0;
    const int k0 = ks * slice, k1 = min(K, k0 + slice);
    const int row = PAIR ? (col / 16) * 32 + (col % 16) : col;
    const bf16* w0 = W + (long long)row * ldw;
    for (int m0 = 0; m0 < B; m0 += DS_MR) {
      float acc[DS_MR], accg[DS_MR];
#pragma unroll
      for (int b = 0; b < DS_MR; ++b) acc[b] = accg[b] = 0.f;
      if (active) {
#pragma unroll
        for (int i = 0; i < DS_MAXC; ++i) {  // chunks whose weights were requested before the barrier (round 0)
          const int c = k0 + (lane + 32 * i) * 8;
          if (c < k1) {
            uint4 wv = wp.v[i], wg = wp.g[i];
            if (r != 0) {
              wv = __ldg(reinterpret_cast<const uint4*>(w0 + c));
              if (PAIR) wg = __ldg(reinterpret_cast<const uint4*>(w0 + (long long)16 * ldw + c));
            }
#pragma unroll
            for (int b = 0; b < DS_MR; ++b)
              if (m0 + b < B) {
                const uint4 av = *reinterpret_cast<const uint4*>(As + (m0 + b) * lda_s + c);
                acc[b] += dot8f(av, wv);
                if (PAIR) accg[b] += dot8f(av, wg);
              }
          }
        }
        for (int c = k0 + (lane + 32 * DS_MAXC) * 8; c < k1; c += 256) {  // longer slices: stream the rest
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w0 + c));
          uint4 wg = make_uint4(0u, 0u, 0u, 0u);
          if (PAIR) wg = __ldg(reinterpret_cast<const uint4*>(w0 + (long long)16 * ldw + c));
#pragma unroll
          for (int b = 0; b < DS_MR; ++b)
            if (m0 + b < B) {
              const uint4 av = *reinterpret_cast<const uint4*>(As + (m0 + b) * lda_s + c);
              acc[b] += dot8f(av, wv);
              if (PAIR) accg[b] += dot8f(av, wg);
            }
        }
      }
      float mine = 0.f, mineg = 0.f;
#pragma unroll
      for (int b = 0; b < DS_MR; ++b) {
        const float v = warp_sum(acc[b]);
        const float g = PAIR ? warp_sum(accg[b]) : 0.f;
        if (lane == b) { mine = v; mineg = g; }
      }
      if (S > 1) {
        if (lane < DS_MR) {
          red[(warp * 2 + 0) * DS_MR + lane] = mine;
          red[(warp * 2 + 1) * DS_MR + lane] = mineg;
        }
        __syncthreads();
        if (ks == 0 && lane < DS_MR) {
          for (int s2 = 1; s2 < S; ++s2) {
            mine += red[((warp + s2) * 2 + 0) * DS_MR + lane];
            mineg += red[((warp + s2) * 2 + 1) * DS_MR + lane];
          }
        }
        __syncthreads();
      }
      if (active && ks == 0 && lane < DS_MR && m0 + lane < B) {
        const long long m = m0 + lane;
        float v = mine;
        if (PAIR) {
          v = v * gelu_erf(mineg);
        } else if (o.bias != nullptr) {
          v += __ldg(o.bias + col);
        }
        if (o.f32 != nullptr) o.f32[m * o.ld + col] = v;
        if (o.b16 != nullptr) o.b16[m * o.ld + col] = __float2bfloat16(v);
      }
    }
  }
}

// stage a bf16 [B][K] operand written by other CTAs in the previous phase (global, L1 bypassed) into shared memory
__device__ __forceinline__ void stage_act(bf16* As, int lda_s, const bf16* src, int ld_src, int B, int K) {
  const int k8 = K / 8;
  for (int i = threadIdx.x; i < B * k8; i += DS_THREADS) {
    const int b = i / k8, c = (i - b * k8) * 8;
    *reinterpret_cast<uint4*>(As + b * lda_s + c) = __ldcg(reinterpret_cast<const uint4*>(src + (long long)b * ld_src + c));
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// row norms (one warp per sample row; every CTA does all B rows)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ds_row_stats(const float4 (&v)[DS_LNV], int nv, int D, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DS_LNV; ++i)
    if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
  s = warp_sum(s);
  mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < DS_LNV; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  q = warp_sum(q);
  rstd = rsqrtf(q / (float)D + 1e-5f);
}

// Requested BEFORE the barrier that precedes the norms (cp.async, L2 -> shared memory, no registers held):
// lnp[0..1] = post-norm weight / bias of `prev`, lnp[2..3] = pre-norm weight / bias of `cur` (w2 / b2 when cur is
// NULL: the final StableLayerNorm), shs[b][0:D/2] = the two shifted channel quarters of ShiftVideoTokens, taken from
// the pre-norm rows of positions t - fmap and t - 1 (zeros at the grid border).
__device__ __forceinline__ void prefetch_norms(const DecParams& p, const DecSub* prev, const DecSub* cur, const float* w2,
                                               const float* b2, int t, float* lnp, bf16* shs) {
  const int D = p.D, D4 = D / 4;
  for (int i = threadIdx.x; i < 4 * D4; i += DS_THREADS) {
    const int arr = i / D4, k = i - arr * D4;
    const float* src = arr == 0 ? (prev ? prev->post_w : nullptr)
                     : arr == 1 ? (prev ? prev->post_b : nullptr)
                     : arr == 2 ? (cur ? cur->pre_w : w2)
                                : (cur ? cur->pre_b : b2);
    if (src != nullptr) cp_async16(lnp + arr * D + k * 4, src + k * 4);
  }
  if (cur != nullptr && cur->shift && t >= 1) {
    const int T = p.fmap * p.fmap;
    const int pos = (t - 1) % T;
    const int row = pos / p.fmap, col = pos - row * p.fmap;
    const int src_h = row > 0 ? t - p.fmap : -1, src_w = col > 0 ? t - 1 : -1;
    const int q4 = D / 4, qp = D / 32;  // 16-byte pieces per channel quarter (bf16)
    const bf16* sc = reinterpret_cast<const bf16*>(cur->shift_cache);
    for (int i = threadIdx.x; i < p.B * 2 * qp; i += DS_THREADS) {
      const int b = i / (2 * qp), r = i - b * 2 * qp;
      const int quarter = r / qp, k = r - quarter * qp;
      const int src = quarter == 0 ? src_h : src_w;
      bf16* dst = shs + b * (D / 2) + quarter * q4 + k * 8;
      if (src >= 0) cp_async16(dst, sc + ((long long)b * p.npos + src) * D + quarter * q4 + k * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  cp_async_commit();
}

// prev != NULL:  streams[prev->write] += LayerNorm_post(y)        (SandwichNorm tail + residual)
// cur  != NULL:  As = bf16(LayerNorm_pre(streams[cur->read])) with the ShiftVideoTokens gather
// Needs prefetch_norms(prev, cur) to have been issued; waits for it here.
__device__ __forceinline__ void ln_prologue(const DecParams& p, const DecSub* prev, const DecSub* cur, int t, float* streams,
                                            bf16* As, int lda_s, const float* lnp, const bf16* shs, int warp, int lane) {
  const int D = p.D, B = p.B;
  float4 v[DS_LNV];
  int nv = 0;
#pragma unroll
  for (int i = 0; i < DS_LNV; ++i)
    if ((lane + 32 * i) * 4 < D) nv = i + 1;
  if (warp < B && prev != nullptr) {  // the only loads that had to wait for the barrier
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i)
      if (i < nv) v[i] = __ldcg(reinterpret_cast<const float4*>(p.y + (long long)warp * D + (lane + 32 * i) * 4));
  }
  cp_async_wait_all();
  __syncthreads();
  if (warp < B) {
    const int b = warp;
    if (prev != nullptr) {
      float* st = streams + ((long long)prev->write * B + b) * D;
      float mean, rstd;
      ds_row_stats(v, nv, D, mean, rstd);
#pragma unroll
      for (int i = 0; i < DS_LNV; ++i)
        if (i < nv) {
          const int c = (lane + 32 * i) * 4;
          const float4 w = *reinterpret_cast<const float4*>(lnp + c);
          const float4 bb = *reinterpret_cast<const float4*>(lnp + D + c);
          const float4 r = *reinterpret_cast<const float4*>(st + c);
          v[i].x = (v[i].x - mean) * rstd * w.x + bb.x + r.x;
          v[i].y = (v[i].y - mean) * rstd * w.y + bb.y + r.y;
          v[i].z = (v[i].z - mean) * rstd * w.z + bb.z + r.z;
          v[i].w = (v[i].w - mean) * rstd * w.w + bb.w + r.w;
          *reinterpret_cast<float4*>(st + c) = v[i];
        }
    }
    if (cur != nullptr) {
      if (prev == nullptr || prev->write != cur->read) {
        const float* st = streams + ((long long)cur->read * B + b) * D;
#pragma unroll
        for (int i = 0; i < DS_LNV; ++i)
          if (i < nv) v[i] = *reinterpret_cast<const float4*>(st + (lane + 32 * i) * 4);
      }
      float mean, rstd;
      ds_row_stats(v, nv, D, mean, rstd);
      const int q4 = D / 4;
      const bool shifted = cur->shift && t >= 1;
      bf16* sc = cur->shift ? reinterpret_cast<bf16*>(cur->shift_cache) + (long long)b * p.npos * D : nullptr;
#pragma unroll
      for (int i = 0; i < DS_LNV; ++i)
        if (i < nv) {
          const int c = (lane + 32 * i) * 4;
          const float4 w = *reinterpret_cast<const float4*>(lnp + 2 * D + c);
          const float4 bb = *reinterpret_cast<const float4*>(lnp + 3 * D + c);
          uint2 pk;
          pk.x = pack_bf16x2((v[i].x - mean) * rstd * w.x + bb.x, (v[i].y - mean) * rstd * w.y + bb.y);
          pk.y = pack_bf16x2((v[i].z - mean) * rstd * w.z + bb.z, (v[i].w - mean) * rstd * w.w + bb.w);
          uint2 outv = pk;
          if (cur->shift && c < 2 * q4) {
            // every CTA holds the same value; CTA 0 publishes it for the tokens to come
            if (blockIdx.x == 0 && t < p.npos) *reinterpret_cast<uint2*>(sc + (long long)t * D + c) = pk;
            if (shifted) outv = *reinterpret_cast<const uint2*>(shs + b * (D / 2) + c);
          }
          *reinterpret_cast<uint2*>(As + b * lda_s + c) = outv;
        }
    }
  }
  __syncthreads();
}

// final StableLayerNorm of streams[0] (+ streams[1]) -> out_f32 / out_bf16 (CTA 0) and As (every CTA, for the logits);
// weight / bias in lnp[2..3]
__device__ __forceinline__ void stable_ln_rows(const DecParams& p, const float* streams, bf16* As, int lda_s,
                                               const float* lnp, int warp, int lane) {
  const int D = p.D, B = p.B;
  if (warp < B) {
    const int b = warp;
    float4 v[DS_LNV];
    int nv = 0;
    float mx = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < D) {
        nv = i + 1;
        v[i] = *reinterpret_cast<const float4*>(streams + (long long)b * D + c);
        if (p.reversible) {
          const float4 u = *reinterpret_cast<const float4*>(streams + ((long long)B + b) * D + c);
          v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
        }
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
    }
    mx = warp_max(mx);
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i)
      if (i < nv) { v[i].x /= mx; v[i].y /= mx; v[i].z /= mx; v[i].w /= mx; }
    float mean, rstd;
    ds_row_stats(v, nv, D, mean, rstd);
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        const float4 ww = *reinterpret_cast<const float4*>(lnp + 2 * D + c);
        const float4 bb = *reinterpret_cast<const float4*>(lnp + 3 * D + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
        o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
        o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
        o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(As + b * lda_s + c) = pk;
        if (blockIdx.x == 0) {
          if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + (long long)b * D + c) = o;
          if (p.out_bf16 != nullptr) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out_bf16) + (long long)b * D + c) = pk;
        }
      }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// attention pieces (one query row per sample)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int key3dna(const DecParams& p, const DecSub& s, int t, int nv, int j, int& row) {
  if (j == 0) { row = 0; return DK_NORMAL; }
  const int jj = j - 1;
  const int c = jj % s.kw, bq = (jj / s.kw) % s.kh, a = jj / (s.kw * s.kh);
  const int T = p.fmap * p.fmap;
  const int vt = t - 1;
  const int f = vt / T, y = (vt % T) / p.fmap, x = vt % p.fmap;
  const int pf = s.dt * (s.kt - 1) / 2, ph = s.dh_ * (s.kh - 1) / 2, pw = s.dw * (s.kw - 1) / 2;
  const int Pf = p.causal ? 2 * pf : pf, Ph = p.causal ? 2 * ph : ph, Pw = p.causal ? 2 * pw : pw;
  const int ff = f + a * s.dt - Pf, yy = y + bq * s.dh_ - Ph, xx = x + c * s.dw - Pw;
  if (ff < 0 || ff >= p.max_frames || yy < 0 || yy >= p.fmap || xx < 0 || xx >= p.fmap) return DK_MASKED;
  const int idx = (ff * p.fmap + yy) * p.fmap + xx;
  if (idx >= nv) return DK_ZERO;
  row = 1 + idx;
  return DK_NORMAL;
}

__device__ __forceinline__ float dot8q(const float* q, const uint4& u) {
  const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  const float4 q0 = *reinterpret_cast<const float4*>(q), q1 = *reinterpret_cast<const float4*>(q + 4);
  return q0.x * a.x + q0.y * a.y + q0.z * b2.x + q0.w * b2.y + q1.x * c.x + q1.y * c.y + q1.z * d.x + q1.w * d.y;
}

// scores of heads [0, nh) (pointers already offset to the first head): S[hl*J + j].  DH8 = dh / 8 as a compile-time
// constant puts every 16-byte load of a key row in flight at once (0 = runtime head width).
template <int DH8>
__device__ __forceinline__ void attn_scores(const float* qs, const int* keys, const bf16* kbase, int k_rs,
                                            const float* null_k, int nh, int dh, int J, float* S) {
  for (int item = threadIdx.x; item < nh * J; item += DS_THREADS) {
    const int hl = item / J, j = item - hl * J;
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    const float* q = qs + hl * dh;
    float s;
    if (kind == DK_NORMAL) {
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (long long)row * k_rs + hl * dh);
      s = 0.f;
      if (DH8 > 0) {
        uint4 u[DH8 > 0 ? DH8 : 1];
#pragma unroll
        for (int i = 0; i < DH8; ++i) u[i] = __ldcg(kr + i);
#pragma unroll
        for (int i = 0; i < DH8; ++i) s += dot8q(q + i * 8, u[i]);
      } else {
        for (int i = 0; i < dh / 8; ++i) s += dot8q(q + i * 8, __ldcg(kr + i));
      }
    } else if (kind == DK_NULL) {
      const float* nk = null_k + hl * dh;
      s = 0.f;
      for (int i = 0; i < dh; ++i) s += q[i] * __ldg(nk + i);
    } else {
      s = (kind == DK_MASKED) ? -FLT_MAX : 0.f;
    }
    S[item] = s;
  }
}

// in-place fp32 softmax of rows [0, nh) of S (warp per row)
__device__ __forceinline__ void attn_softmax(float* S, int nh, int J, int warp, int lane) {
  for (int hl = warp; hl < nh; hl += DS_WARPS) {
    float* Sw = S + hl * J;
    float m = -FLT_MAX;
    for (int j = lane; j < J; j += 32) m = fmaxf(m, Sw[j]);
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < J; j += 32) {
      const float e = __expf(Sw[j] - m);
      Sw[j] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.0f / l;
    for (int j = lane; j < J; j += 32) Sw[j] *= inv;
  }
}

// out[hl*dh + c] = sum_j P[hl*J + j] * V[row_j][hl*dh + c]   for heads [0, nh) (vbase / null_v offset to the first head).
// Branch-free key loop (masked / zero / null keys load row 0 with weight 0) so that the unrolled loads overlap.
__device__ __forceinline__ void attn_pv(const float* P, const int* keys, const bf16* vbase, int v_rs, const float* null_v,
                                        int nh, int dh, int J, float* part, float* outs) {
  const int npairs = nh * dh / 2;
  const int KG = DS_THREADS / npairs;  // key groups (>= 1: H*dh <= 1024)
  const int cp = threadIdx.x % npairs, kg = threadIdx.x / npairs;
  if (kg < KG) {
    const int hl = cp / (dh / 2), c2 = cp - hl * (dh / 2);
    const float* Ph = P + hl * J;
    const bf16* vcol = vbase + hl * dh + 2 * c2;
    float ax = 0.f, ay = 0.f;
    int j = kg;
    for (; j + 7 * KG < J; j += 8 * KG) {
      uint32_t u[8];
      float pj[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kj = keys[j + k * KG];
        const bool normal = (kj >> 28) == DK_NORMAL;
        u[k] = __ldcg(reinterpret_cast<const unsigned int*>(vcol + (long long)(normal ? (kj & 0x0FFFFFFF) : 0) * v_rs));
        pj[k] = normal ? Ph[j + k * KG] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 v = unpack_bf16x2(u[k]);
        ax = fmaf(pj[k], v.x, ax);
        ay = fmaf(pj[k], v.y, ay);
      }
    }
    for (; j < J; j += KG) {
      const int kj = keys[j];
      if ((kj >> 28) == DK_NORMAL) {
        const float2 v = unpack_bf16x2(__ldcg(reinterpret_cast<const unsigned int*>(vcol + (long long)(kj & 0x0FFFFFFF) * v_rs)));
        ax = fmaf(Ph[j], v.x, ax);
        ay = fmaf(Ph[j], v.y, ay);
      }
    }
    if (kg == 0 && (keys[0] >> 28) == DK_NULL) {  // the learned null key / value is slot 0
      ax = fmaf(Ph[0], __ldg(null_v + hl * dh + 2 * c2), ax);
      ay = fmaf(Ph[0], __ldg(null_v + hl * dh + 2 * c2 + 1), ay);
    }
    part[(kg * npairs + cp) * 2 + 0] = ax;
    part[(kg * npairs + cp) * 2 + 1] = ay;
  }
  __syncthreads();
  if (threadIdx.x < npairs) {
    float ax = 0.f, ay = 0.f;
    for (int g = 0; g < KG; ++g) {
      ax += part[(g * npairs + threadIdx.x) * 2 + 0];
      ay += part[(g * npairs + threadIdx.x) * 2 + 1];
    }
    outs[2 * threadIdx.x] = ax;
    outs[2 * threadIdx.x + 1] = ay;
  }
  __syncthreads();
}

__device__ __forceinline__ void attn_scores_any(const float* qs, const int* keys, const bf16* kbase, int k_rs,
                                                const float* null_k, int nh, int dh, int J, float* S) {
  if (dh == 64) attn_scores<8>(qs, keys, kbase, k_rs, null_k, nh, dh, J, S);
  else if (dh == 32) attn_scores<4>(qs, keys, kbase, k_rs, null_k, nh, dh, J, S);
  else attn_scores<0>(qs, keys, kbase, k_rs, null_k, nh, dh, J, S);
}

// key list of the dense cross attention: slot 0 = learned null key, slot 1 + i = context token i (masked by key_mask)
__device__ __forceinline__ void cross_keys(const DecParams& p, int b, int J, int* keys) {
  for (int j = threadIdx.x; j < J; j += DS_THREADS) {
    int kd = DK_NULL, row = 0;
    if (j > 0) {
      row = j - 1;
      kd = (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + row] == 0) ? DK_MASKED : DK_NORMAL;
    }
    keys[j] = (kd << 28) | row;
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
#define DS_STAMP(id)                                                            \
  do {                                                                          \
    if (p.prof != nullptr && blockIdx.x == p.prof_cta && threadIdx.x == 0) {   \
      p.prof[2 * nstamp] = (long long)(si * 16 + (id));                         \
      p.prof[2 * nstamp + 1] = clock64();                                       \
      ++nstamp;                                                                 \
    }                                                                           \
  } while (0)

__global__ void __launch_bounds__(DS_THREADS, 1) decode_stack_kernel(const DecParams p) {
  extern __shared__ __align__(16) uint8_t ds_smem[];
  const int B = p.B, D = p.D, H = p.H, dh = p.dh, inner = H * dh;
  const DsLayout L = ds_layout(B, D, p.kmax, H, dh, p.jmax);
  float* streams = reinterpret_cast<float*>(ds_smem + L.streams);
  bf16* As = reinterpret_cast<bf16*>(ds_smem + L.act);
  float* red = reinterpret_cast<float*>(ds_smem + L.red);
  float* Ss = reinterpret_cast<float*>(ds_smem + L.S);
  float* Pm = reinterpret_cast<float*>(ds_smem + L.pm);
  int* keys = reinterpret_cast<int*>(ds_smem + L.keys);
  float* qs = reinterpret_cast<float*>(ds_smem + L.qs);
  float* Wt = reinterpret_cast<float*>(ds_smem + L.wt);
  float* part = reinterpret_cast<float*>(ds_smem + L.part);
  float* outs = reinterpret_cast<float*>(ds_smem + L.outs);
  float* lnp = reinterpret_cast<float*>(ds_smem + L.lnp);
  bf16* shs = reinterpret_cast<bf16*>(ds_smem + L.shs);
  constexpr int DESC_STRIDE = (sizeof(DecSub) + 15) & ~15;
  constexpr int DESC_WORDS = sizeof(DecSub) / 4;
  const int lda_s = p.kmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * DS_WARPS + warp;
  const int total_warps = gridDim.x * DS_WARPS;
  const int t = __ldg(p.t_ptr);
  const float qscale = rsqrtf((float)dh);
  const int dbg = p.debug_flags;
  GridBar bar{p.barrier, gridDim.x, 0u};
  bf16* act = reinterpret_cast<bf16*>(p.act);
  bf16* actq = reinterpret_cast<bf16*>(p.actq);
  int nstamp = 0;
  int si = 0;

  // X = x (plain) or [x, x] (reversible.py:133); descriptor of the first sub-block
  for (int i = threadIdx.x; i < B * D; i += DS_THREADS) {
    const float v = __ldg(p.x_in + i);
    streams[i] = v;
    if (p.reversible) streams[B * D + i] = v;
  }
  if (threadIdx.x < DESC_WORDS)
    reinterpret_cast<uint32_t*>(ds_smem + L.desc)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(p.subs) + threadIdx.x);
  __syncthreads();
  DS_STAMP(0);

  const DecSub* prev = nullptr;
  for (si = 0; si < p.nsubs; ++si) {
    const DecSub* s = reinterpret_cast<const DecSub*>(ds_smem + L.desc + (si % 3) * DESC_STRIDE);
    // descriptor of the NEXT sub-block: requested now, parked in shared memory after the first barrier wait
    uint32_t next_word = 0;
    const bool has_next = si + 1 < p.nsubs && threadIdx.x < DESC_WORDS;
    if (has_next) next_word = __ldg(reinterpret_cast<const uint32_t*>(p.subs + si + 1) + threadIdx.x);
    const int kind = s->kind;
    WPre wp;
    // ---- phase 1: (post-norm + residual of the previous sub-block,) pre-norm, first product ----
    int N1, S1;
    if (kind == NUWA_DEC_3DNA) { N1 = 3 * inner; S1 = 1; }
    else if (kind == NUWA_DEC_CROSS) { N1 = inner; S1 = p.split_small; }
    else { N1 = s->ip; S1 = 1; }
    const bf16* Wa = reinterpret_cast<const bf16*>(s->w_a);
    const bf16* Wb = reinterpret_cast<const bf16*>(s->w_b);
    if (kind == NUWA_DEC_FF) gemv_prefetch<true>(wp, Wa, D, N1, D, S1, gwarp, lane);
    else gemv_prefetch<false>(wp, Wa, D, N1, D, S1, gwarp, lane);
    prefetch_norms(p, prev, s, nullptr, nullptr, t, lnp, shs);
    DS_STAMP(1);
    if (prev != nullptr) grid_barrier(bar);  // y of the previous sub-block is complete
    DS_STAMP(2);
    if (has_next)
      reinterpret_cast<uint32_t*>(ds_smem + L.desc + ((si + 1) % 3) * DESC_STRIDE)[threadIdx.x] = next_word;
    if (!(dbg & 1)) ln_prologue(p, prev, s, t, streams, As, lda_s, lnp, shs, warp, lane);
    else { cp_async_wait_all(); __syncthreads(); }
    DS_STAMP(3);
    if (kind == NUWA_DEC_3DNA) {
      // q|k|v of the new token go straight into row t of the cache
      bf16* cache = reinterpret_cast<bf16*>(s->cache);
      GemvOut o{nullptr, cache + (long long)t * 3 * inner, (long long)p.npos * 3 * inner, nullptr};
      if (!(dbg & 2)) gemv_run<false>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, inner, D, inner, p.split_small, gwarp, lane);
      for (int i = threadIdx.x; i < H * H; i += DS_THREADS) Wt[i] = __ldg(s->talk + i);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      // ---- phase 2: attention, one CTA per sample (all heads: talking heads mix across heads) ----
      if (!(dbg & 4))
      for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const bf16* cb = cache + (long long)b * p.npos * 3 * inner;
        if (t == 0) {  // bos attends only to itself (:499,608)
          for (int c = threadIdx.x; c < inner; c += DS_THREADS) act[(long long)b * inner + c] = __ldcg(cb + 2 * inner + c);
          continue;
        }
        const int J = 1 + s->kt * s->kh * s->kw;
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {
          int row = 0;
          const int kd = key3dna(p, *s, t, t, j, row);
          keys[j] = (kd << 28) | row;
        }
        for (int c = threadIdx.x; c < inner; c += DS_THREADS)
          qs[c] = __bfloat162float(__ldcg(cb + (long long)t * 3 * inner + c)) * qscale;
        __syncthreads();
        attn_scores_any(qs, keys, cb + inner, 3 * inner, nullptr, H, dh, J, Ss);
        __syncthreads();
        attn_softmax(Ss, H, J, warp, lane);
        __syncthreads();
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {  // talking heads (:556-558)
          float pin[16];
          for (int h = 0; h < H; ++h) pin[h] = Ss[h * J + j];
          for (int g = 0; g < H; ++g) {
            float a = 0.f;
            for (int h = 0; h < H; ++h) a = fmaf(Wt[g * H + h], pin[h], a);
            Ss[g * J + j] = a;
          }
        }
        __syncthreads();
        attn_pv(Ss, keys, cb + 2 * inner, 3 * inner, nullptr, H, dh, J, part, outs);
        for (int c = threadIdx.x; c < inner; c += DS_THREADS) act[(long long)b * inner + c] = __float2bfloat16(outs[c]);
        __syncthreads();
      }
      DS_STAMP(6);
      grid_barrier(bar);
      DS_STAMP(7);
      // ---- phase 3: to_out (+ bias) ----
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, s->b_out};
      if (!(dbg & 2)) gemv_run<false>(wp, Wb, inner, D, inner, p.split_small, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    } else if (kind == NUWA_DEC_CROSS) {
      GemvOut o{nullptr, actq, (long long)inner, nullptr};
      if (!(dbg & 2)) gemv_run<false>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, inner, D, inner, p.split_small, gwarp, lane);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      const int J = p.nk + 1;
      const bf16* kv = reinterpret_cast<const bf16*>(s->cache);
      // ---- phase 2a: scores of one (sample, head) per CTA ----
      if (!(dbg & 4))
      for (int w = blockIdx.x; w < B * H; w += gridDim.x) {
        const int b = w / H, h = w - b * H;
        cross_keys(p, b, J, keys);
        for (int c = threadIdx.x; c < dh; c += DS_THREADS)
          qs[c] = __bfloat162float(__ldcg(actq + (long long)b * inner + h * dh + c)) * qscale;
        __syncthreads();
        attn_scores_any(qs, keys, kv + (long long)b * p.nk * 2 * inner + h * dh, 2 * inner, s->null_k + h * dh, 1, dh, J, Ss);
        __syncthreads();
        for (int j = threadIdx.x; j < J; j += DS_THREADS) p.scores[((long long)b * H + h) * J + j] = Ss[j];
        __syncthreads();
      }
      DS_STAMP(6);
      grid_barrier(bar);
      DS_STAMP(7);
      // ---- phase 2b: softmax of every head of the sample, talking-heads row g, PV of head g ----
      if (!(dbg & 4))
      for (int w = blockIdx.x; w < B * H; w += gridDim.x) {
        const int b = w / H, g = w - b * H;
        cross_keys(p, b, J, keys);
        for (int i = threadIdx.x; i < H * J; i += DS_THREADS) Ss[i] = __ldcg(p.scores + (long long)b * H * J + i);
        for (int i = threadIdx.x; i < H; i += DS_THREADS) Wt[i] = __ldg(s->talk + g * H + i);
        __syncthreads();
        attn_softmax(Ss, H, J, warp, lane);
        __syncthreads();
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {  // :372
          float a = 0.f;
          for (int h = 0; h < H; ++h) a = fmaf(Wt[h], Ss[h * J + j], a);
          Pm[j] = a;
        }
        __syncthreads();
        attn_pv(Pm, keys, kv + (long long)b * p.nk * 2 * inner + inner + g * dh, 2 * inner, s->null_v + g * dh, 1, dh, J,
                part, outs);
        for (int c = threadIdx.x; c < dh; c += DS_THREADS) act[(long long)b * inner + g * dh + c] = __float2bfloat16(outs[c]);
        __syncthreads();
      }
      DS_STAMP(8);
      grid_barrier(bar);
      DS_STAMP(9);
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      if (!(dbg & 2)) gemv_run<false>(wp, Wb, inner, D, inner, p.split_small, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    } else {
      // GEGLU product: value/gate pairs -> a * gelu(g) (:255-258)
      GemvOut o{nullptr, act, (long long)s->ip, nullptr};
      if (!(dbg & 2)) gemv_run<true>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, s->ip, D, s->ip, p.split_ff, gwarp, lane);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      stage_act(As, lda_s, act, s->ip, B, s->ip);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      if (!(dbg & 2)) gemv_run<false>(wp, Wb, s->ip, D, s->ip, p.split_ff, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    }
    prev = s;
  }
  // ---- tail: last post-norm + residual, StableLayerNorm, logits ----
  WPre wl;
  const bf16* Wl = reinterpret_cast<const bf16*>(p.w_logits);
  if (Wl != nullptr) gemv_prefetch<false>(wl, Wl, D, p.V, D, 1, gwarp, lane);
  prefetch_norms(p, prev, nullptr, p.norm_w, p.norm_b, t, lnp, shs);
  grid_barrier(bar);
  DS_STAMP(13);
  ln_prologue(p, prev, nullptr, t, streams, As, lda_s, lnp, shs, warp, lane);
  stable_ln_rows(p, streams, As, lda_s, lnp, warp, lane);
  if (Wl != nullptr) {
    GemvOut ol{p.logits, nullptr, (long long)p.V, nullptr};
    gemv_run<false>(wl, Wl, D, p.V, D, 1, As, lda_s, B, red, ol, gwarp, total_warps, warp, lane);
  }
  DS_STAMP(14);
  // ---- leave: the last CTA out resets the barrier words for the next launch ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prevc = atomicAdd(p.barrier + 1, 1u);
    if (prevc == gridDim.x - 1) {
      p.barrier[0] = 0u;
      p.barrier[1] = 0u;
      __threadfence();
    }
    if (p.prof != nullptr && blockIdx.x == p.prof_cta) p.prof[2 * nstamp] = -1;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int decode_stack(const DecParams& p_in, int cooperative, cudaStream_t stream) {
  DecParams p = p_in;
  if (p.subs == nullptr || p.nsubs <= 0 || p.B <= 0 || p.B > 16 || p.t_ptr == nullptr || p.barrier == nullptr)
    return NUWA_ERR_INVALID;
  if (p.D % 32 != 0 || p.D > 1024 || p.H <= 0 || p.H > 16 || p.dh % 8 != 0 || p.H * p.dh > 1024 || (p.H * p.dh) % 8 != 0)
    return NUWA_ERR_INVALID;
  if (p.kmax < p.D || p.kmax < p.H * p.dh || p.kmax % 8 != 0) return NUWA_ERR_INVALID;
  if (p.x_in == nullptr || p.y == nullptr || p.act == nullptr || p.actq == nullptr || p.norm_w == nullptr ||
      p.norm_b == nullptr)
    return NUWA_ERR_INVALID;
  if (p.w_logits != nullptr && (p.logits == nullptr || p.V <= 0)) return NUWA_ERR_INVALID;
  if (p.j3max < 1 || p.j3max > 4096) return NUWA_ERR_INVALID;
  p.jmax = p.j3max > p.nk + 1 ? p.j3max : p.nk + 1;
  if (p.nk > 0 && p.scores == nullptr) return NUWA_ERR_INVALID;
  if (p.split_small != 1 && p.split_small != 2 && p.split_small != 4) p.split_small = 2;
  if (p.split_ff != 1 && p.split_ff != 2 && p.split_ff != 4) p.split_ff = 4;
  const DsLayout L = ds_layout(p.B, p.D, p.kmax, p.H, p.dh, p.jmax);
  if (L.total > 200 * 1024) return NUWA_ERR_INVALID;
  static int max_blocks_per_sm_smem = -1, cached_smem = -1;
  if (cudaFuncSetAttribute(decode_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess)
    return NUWA_ERR_CUDA;
  if (cached_smem != L.total) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_stack_kernel, DS_THREADS, (size_t)L.total) != cudaSuccess)
      return NUWA_ERR_CUDA;
    max_blocks_per_sm_smem = nb;
    cached_smem = L.total;
  }
  if (max_blocks_per_sm_smem < 1) return NUWA_ERR_INVALID;
  int grid = device_sm_count();
  if (p.max_ctas > 0 && p.max_ctas < grid) grid = p.max_ctas;
  if (grid < 1) return NUWA_ERR_CUDA;
  if (cooperative) {
    void* args[] = {(void*)&p};
    if (cudaLaunchCooperativeKernel((const void*)decode_stack_kernel, dim3(grid), dim3(DS_THREADS), args, (size_t)L.total,
                                    stream) != cudaSuccess) {
      cudaGetLastError();
      return NUWA_ERR_CUDA;
    }
    ++g_launch_count;
    return NUWA_OK;
  }
  // plain launch: grid <= SM count and one CTA per SM, so all CTAs become resident once earlier work drains
  decode_stack_kernel<<<grid, DS_THREADS, L.total, stream>>>(p);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
