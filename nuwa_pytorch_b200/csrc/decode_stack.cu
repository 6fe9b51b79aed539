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
  int streams, act, red, S, pm, keys, qs, wt, part, outs, total;
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
    __threadfence();
    atomicAdd(g.count, 1u);
    long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_u32(g.count) < g.target) {
      ++spins;
      if (spins == 4096u) t0 = clock64();
      // watchdog: a protocol bug becomes a launch error instead of a hung GPU (~2 s)
      if (spins > 4096u && (spins & 4095u) == 0u && (clock64() - t0) > 4000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

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

// prev != NULL:  streams[prev->write] += LayerNorm_post(y)        (SandwichNorm tail + residual)
// cur  != NULL:  As = bf16(LayerNorm_pre(streams[cur->read])) with the ShiftVideoTokens gather
__device__ __forceinline__ void ln_prologue(const DecParams& p, const DecSub* prev, const DecSub* cur, int t, float* streams,
                                            bf16* As, int lda_s, int warp, int lane) {
  const int D = p.D, B = p.B;
  if (warp < B) {
    const int b = warp;
    int nv = 0;
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i)
      if ((lane + 32 * i) * 4 < D) nv = i + 1;
    float4 v[DS_LNV];
    if (prev != nullptr) {
      float* st = streams + ((long long)prev->write * B + b) * D;
#pragma unroll
      for (int i = 0; i < DS_LNV; ++i)
        if (i < nv) v[i] = __ldcg(reinterpret_cast<const float4*>(p.y + (long long)b * D + (lane + 32 * i) * 4));
      float mean, rstd;
      ds_row_stats(v, nv, D, mean, rstd);
#pragma unroll
      for (int i = 0; i < DS_LNV; ++i)
        if (i < nv) {
          const int c = (lane + 32 * i) * 4;
          const float4 w = __ldg(reinterpret_cast<const float4*>(prev->post_w + c));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(prev->post_b + c));
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
      int src_h = -1, src_w = -1;
      if (shifted) {
        const int T = p.fmap * p.fmap;
        const int pos = (t - 1) % T;
        const int row = pos / p.fmap, col = pos - row * p.fmap;
        if (row > 0) src_h = t - p.fmap;
        if (col > 0) src_w = t - 1;
      }
      bf16* sc = cur->shift ? reinterpret_cast<bf16*>(cur->shift_cache) + (long long)b * p.npos * D : nullptr;
#pragma unroll
      for (int i = 0; i < DS_LNV; ++i)
        if (i < nv) {
          const int c = (lane + 32 * i) * 4;
          const float4 w = __ldg(reinterpret_cast<const float4*>(cur->pre_w + c));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(cur->pre_b + c));
          uint2 pk;
          pk.x = pack_bf16x2((v[i].x - mean) * rstd * w.x + bb.x, (v[i].y - mean) * rstd * w.y + bb.y);
          pk.y = pack_bf16x2((v[i].z - mean) * rstd * w.z + bb.z, (v[i].w - mean) * rstd * w.w + bb.w);
          uint2 outv = pk;
          if (cur->shift && c < 2 * q4) {
            // every CTA holds the same value; CTA 0 publishes it for the tokens to come
            if (blockIdx.x == 0 && t < p.npos) *reinterpret_cast<uint2*>(sc + (long long)t * D + c) = pk;
            if (shifted) {
              const int src = c < q4 ? src_h : src_w;
              outv = src >= 0 ? *reinterpret_cast<const uint2*>(sc + (long long)src * D + c) : make_uint2(0u, 0u);
            }
          }
          *reinterpret_cast<uint2*>(As + b * lda_s + c) = outv;
        }
    }
  }
  __syncthreads();
}

// final StableLayerNorm of streams[0] (+ streams[1]) -> out_f32 / out_bf16 (CTA 0) and As (every CTA, for the logits)
__device__ __forceinline__ void stable_ln_rows(const DecParams& p, const float* streams, bf16* As, int lda_s, int warp,
                                               int lane) {
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
        const float4 ww = __ldg(reinterpret_cast<const float4*>(p.norm_w + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.norm_b + c));
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
__device__ __forceinline__ int key3dna(const DecParams& p, int t, int nv, int j, int& row) {
  if (j == 0) { row = 0; return DK_NORMAL; }
  const int jj = j - 1;
  const int c = jj % p.kw, bq = (jj / p.kw) % p.kh, a = jj / (p.kw * p.kh);
  const int T = p.fmap * p.fmap;
  const int vt = t - 1;
  const int f = vt / T, y = (vt % T) / p.fmap, x = vt % p.fmap;
  const int pf = p.dt * (p.kt - 1) / 2, ph = p.dh_ * (p.kh - 1) / 2, pw = p.dw * (p.kw - 1) / 2;
  const int Pf = p.causal ? 2 * pf : pf, Ph = p.causal ? 2 * ph : ph, Pw = p.causal ? 2 * pw : pw;
  const int ff = f + a * p.dt - Pf, yy = y + bq * p.dh_ - Ph, xx = x + c * p.dw - Pw;
  if (ff < 0 || ff >= p.max_frames || yy < 0 || yy >= p.fmap || xx < 0 || xx >= p.fmap) return DK_MASKED;
  const int idx = (ff * p.fmap + yy) * p.fmap + xx;
  if (idx >= nv) return DK_ZERO;
  row = 1 + idx;
  return DK_NORMAL;
}

// scores of heads [0, nh) (pointers already offset to the first head): S[hl*J + j]
__device__ __forceinline__ void attn_scores(const float* qs, const int* keys, const bf16* kbase, int k_rs,
                                            const float* null_k, int nh, int dh, int J, float* S) {
  for (int item = threadIdx.x; item < nh * J; item += DS_THREADS) {
    const int hl = item / J, j = item - hl * J;
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    const float* q = qs + hl * dh;
    float s;
    if (kind == DK_NORMAL) {
      const bf16* kr = kbase + (long long)row * k_rs + hl * dh;
      s = 0.f;
      for (int i = 0; i < dh / 8; ++i) {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(kr) + i);
        const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        const float4 q0 = *reinterpret_cast<const float4*>(q + i * 8), q1 = *reinterpret_cast<const float4*>(q + i * 8 + 4);
        s += q0.x * a.x + q0.y * a.y + q0.z * b2.x + q0.w * b2.y + q1.x * c.x + q1.y * c.y + q1.z * d.x + q1.w * d.y;
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

// out[hl*dh + c] = sum_j P[hl*J + j] * V[row_j][hl*dh + c]   for heads [0, nh) (vbase / null_v offset to the first head)
__device__ __forceinline__ void attn_pv(const float* P, const int* keys, const bf16* vbase, int v_rs, const float* null_v,
                                        int nh, int dh, int J, float* part, float* outs) {
  const int npairs = nh * dh / 2;
  const int KG = DS_THREADS / npairs;  // key groups (>= 1: H*dh <= 1024)
  const int cp = threadIdx.x % npairs, kg = threadIdx.x / npairs;
  if (kg < KG) {
    const int hl = cp / (dh / 2), c2 = cp - hl * (dh / 2);
    const float* Ph = P + hl * J;
    float ax = 0.f, ay = 0.f;
#pragma unroll 4
    for (int j = kg; j < J; j += KG) {
      const int kj = keys[j];
      const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
      if (kind == DK_NORMAL) {
        const uint32_t u = __ldcg(reinterpret_cast<const unsigned int*>(vbase + (long long)row * v_rs + hl * dh + 2 * c2));
        const float2 v = unpack_bf16x2(u);
        const float pj = Ph[j];
        ax = fmaf(pj, v.x, ax);
        ay = fmaf(pj, v.y, ay);
      } else if (kind == DK_NULL) {
        const float pj = Ph[j];
        ax = fmaf(pj, __ldg(null_v + hl * dh + 2 * c2), ax);
        ay = fmaf(pj, __ldg(null_v + hl * dh + 2 * c2 + 1), ay);
      }
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

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
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
  const int lda_s = p.kmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * DS_WARPS + warp;
  const int total_warps = gridDim.x * DS_WARPS;
  const int t = __ldg(p.t_ptr);
  const float qscale = rsqrtf((float)dh);
  GridBar bar{p.barrier, gridDim.x, 0u};
  bf16* act = reinterpret_cast<bf16*>(p.act);
  bf16* actq = reinterpret_cast<bf16*>(p.actq);

  // X = x (plain) or [x, x] (reversible.py:133)
  for (int i = threadIdx.x; i < B * D; i += DS_THREADS) {
    const float v = __ldg(p.x_in + i);
    streams[i] = v;
    if (p.reversible) streams[B * D + i] = v;
  }
  __syncthreads();

  const DecSub* prev = nullptr;
  for (int si = 0; si < p.nsubs; ++si) {
    const DecSub* s = p.subs + si;
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
    if (prev != nullptr) grid_barrier(bar);  // y of the previous sub-block is complete
    ln_prologue(p, prev, s, t, streams, As, lda_s, warp, lane);
    if (kind == NUWA_DEC_3DNA) {
      // q|k|v of the new token go straight into row t of the cache
      bf16* cache = reinterpret_cast<bf16*>(s->cache);
      GemvOut o{nullptr, cache + (long long)t * 3 * inner, (long long)p.npos * 3 * inner, nullptr};
      gemv_run<false>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, inner, D, inner, p.split_small, gwarp, lane);
      grid_barrier(bar);
      // ---- phase 2: attention, one CTA per sample (all heads: talking heads mix across heads) ----
      for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const bf16* cb = cache + (long long)b * p.npos * 3 * inner;
        if (t == 0) {  // bos attends only to itself (:499,608)
          for (int c = threadIdx.x; c < inner; c += DS_THREADS) act[(long long)b * inner + c] = __ldcg(cb + 2 * inner + c);
          continue;
        }
        const int J = 1 + p.kt * p.kh * p.kw;
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {
          int row = 0;
          const int kd = key3dna(p, t, t, j, row);
          keys[j] = (kd << 28) | row;
        }
        for (int c = threadIdx.x; c < inner; c += DS_THREADS)
          qs[c] = __bfloat162float(__ldcg(cb + (long long)t * 3 * inner + c)) * qscale;
        for (int i = threadIdx.x; i < H * H; i += DS_THREADS) Wt[i] = __ldg(s->talk + i);
        __syncthreads();
        attn_scores(qs, keys, cb + inner, 3 * inner, nullptr, H, dh, J, Ss);
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
      grid_barrier(bar);
      // ---- phase 3: to_out (+ bias) ----
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, s->b_out};
      gemv_run<false>(wp, Wb, inner, D, inner, p.split_small, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
    } else if (kind == NUWA_DEC_CROSS) {
      GemvOut o{nullptr, actq, (long long)inner, nullptr};
      gemv_run<false>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, inner, D, inner, p.split_small, gwarp, lane);
      grid_barrier(bar);
      const int J = p.nk + 1;
      const bf16* kv = reinterpret_cast<const bf16*>(s->cache);
      // ---- phase 2a: scores of one (sample, head) per CTA ----
      for (int w = blockIdx.x; w < B * H; w += gridDim.x) {
        const int b = w / H, h = w - b * H;
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {
          int kd = DK_NULL, row = 0;
          if (j > 0) {
            row = j - 1;
            kd = (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + row] == 0) ? DK_MASKED : DK_NORMAL;
          }
          keys[j] = (kd << 28) | row;
        }
        for (int c = threadIdx.x; c < dh; c += DS_THREADS)
          qs[c] = __bfloat162float(__ldcg(actq + (long long)b * inner + h * dh + c)) * qscale;
        __syncthreads();
        attn_scores(qs, keys, kv + (long long)b * p.nk * 2 * inner + h * dh, 2 * inner, s->null_k + h * dh, 1, dh, J, Ss);
        __syncthreads();
        for (int j = threadIdx.x; j < J; j += DS_THREADS) p.scores[((long long)b * H + h) * J + j] = Ss[j];
        __syncthreads();
      }
      grid_barrier(bar);
      // ---- phase 2b: softmax of every head of the sample, talking-heads row g, PV of head g ----
      for (int w = blockIdx.x; w < B * H; w += gridDim.x) {
        const int b = w / H, g = w - b * H;
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {
          int kd = DK_NULL, row = 0;
          if (j > 0) {
            row = j - 1;
            kd = (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + row] == 0) ? DK_MASKED : DK_NORMAL;
          }
          keys[j] = (kd << 28) | row;
        }
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
      grid_barrier(bar);
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      gemv_run<false>(wp, Wb, inner, D, inner, p.split_small, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
    } else {
      // GEGLU product: value/gate pairs -> a * gelu(g) (:255-258)
      GemvOut o{nullptr, act, (long long)s->ip, nullptr};
      gemv_run<true>(wp, Wa, D, N1, D, S1, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      gemv_prefetch<false>(wp, Wb, s->ip, D, s->ip, p.split_ff, gwarp, lane);
      grid_barrier(bar);
      stage_act(As, lda_s, act, s->ip, B, s->ip);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      gemv_run<false>(wp, Wb, s->ip, D, s->ip, p.split_ff, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
    }
    prev = s;
  }
  // ---- tail: last post-norm + residual, StableLayerNorm, logits ----
  WPre wl;
  const bf16* Wl = reinterpret_cast<const bf16*>(p.w_logits);
  if (Wl != nullptr) gemv_prefetch<false>(wl, Wl, D, p.V, D, 1, gwarp, lane);
  grid_barrier(bar);
  ln_prologue(p, prev, nullptr, t, streams, As, lda_s, warp, lane);
  stable_ln_rows(p, streams, As, lda_s, warp, lane);
  if (Wl != nullptr) {
    GemvOut ol{p.logits, nullptr, (long long)p.V, nullptr};
    gemv_run<false>(wl, Wl, D, p.V, D, 1, As, lda_s, B, red, ol, gwarp, total_warps, warp, lane);
  }
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
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int decode_stack(const DecParams& p_in, int cooperative, cudaStream_t stream) {
  DecParams p = p_in;
  if (p.subs == nullptr || p.nsubs <= 0 || p.B <= 0 || p.B > 16 || p.t_ptr == nullptr || p.barrier == nullptr)
    return NUWA_ERR_INVALID;
  if (p.D % 16 != 0 || p.D > 1024 || p.H <= 0 || p.H > 16 || p.dh % 8 != 0 || p.H * p.dh > 1024 || (p.H * p.dh) % 8 != 0)
    return NUWA_ERR_INVALID;
  if (p.kmax < p.D || p.kmax < p.H * p.dh || p.kmax % 8 != 0) return NUWA_ERR_INVALID;
  if (p.x_in == nullptr || p.y == nullptr || p.act == nullptr || p.actq == nullptr || p.norm_w == nullptr ||
      p.norm_b == nullptr)
    return NUWA_ERR_INVALID;
  if (p.w_logits != nullptr && (p.logits == nullptr || p.V <= 0)) return NUWA_ERR_INVALID;
  const int j3 = 1 + p.kt * p.kh * p.kw;
  p.jmax = j3 > p.nk + 1 ? j3 : p.nk + 1;
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
