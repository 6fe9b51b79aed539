// Persistent decode-step kernel for generate(): ONE cooperative launch walks a whole decoder stack
// (every SandwichNorm-wrapped Sparse3DNA / cross-attention / FeedForward sub-block, the final
// StableLayerNorm and optionally the to_logits product) for the single new position of every sample.
//
// Why: a decode step of the 64-layer reversible decoder is 256 sub-blocks x ~4 kernels, each a few
// microseconds of work on M = batch rows; as separate launches (even replayed from a CUDA graph) the step
// is bound by launch + dependency latency (~7.6 us per node, 14.6 ms per token), 60x above the time the
// 806 MB of bf16 weights need to stream from HBM.  Here the chip stays resident: 148 CTAs x 16 warps,
// phases separated by a device-wide barrier, and everything that does not depend on the previous phase is
// REQUESTED BEFORE the barrier so that the HBM latency (~1.3 us) hides behind it:
//   * the residual streams live in EVERY CTA's shared memory (each CTA redundantly applies the post-norm +
//     residual + pre-norm of the B rows: no extra barrier, no HBM round trip for the streams),
//   * skinny products (q|k|v, q, GEGLU, out, FF-out, logits) run on the tensor cores: mma.sync m16n8k16 with
//     the 16 weight rows as M and the <= 8 batch rows as N (a CUDA-core version was ISSUE bound: ~1000
//     instructions per output column); a warp owns a (16-row tile, K-slice) task, its weight fragments are
//     16-byte loads issued before the barrier wait (the contraction index is permuted identically for both
//     operands so that a lane's fragment is 8 contiguous bf16), partial tiles are reduced through smem,
//   * K/V rows of earlier tokens (3DNA window) and the context K/V head slices (cross attention) stream into
//     shared memory with cp.async two barriers ahead of their use; norm parameters, to_out bias,
//     talking-heads matrices, shifted channel quarters and the next sub-block's descriptor likewise,
//   * the new token's q|k|v go straight into the KV cache row (no append kernel); cross attention is split
//     over (sample, head) CTAs in two phases (scores | softmax + talking-heads row + PV).
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

static constexpr int DS_THREADS = 256;  // 8 warps, 1 CTA / SM: up to 255 registers per thread (prefetched weight fragments)
static constexpr int DS_WARPS = DS_THREADS / 32;
static constexpr int DS_LNV = 8;      // float4 per lane in the row norms -> D <= 1024
static constexpr int DS_SMEM_MAX = 225 * 1024;
// Warp DS_WWARPS (the last one) is the "sync warp": it signals and polls the device-wide barrier and therefore never
// has loads in flight when it arrives -- a release fence waits for the issuing warp's outstanding loads, so a thread
// that prefetches weights would serialise every barrier behind its own prefetch (HBM latency; measured >= 1 us per
// barrier).  The sync warp takes no product tasks and issues no cp.async / register prefetches; it does share the
// post-barrier work (norm rows, staging, attention), whose loads complete before the next arrival.
static constexpr int DS_WWARPS = DS_WARPS - 1;
static constexpr int DS_WORK = DS_WWARPS * 32;
__device__ __forceinline__ int ds_wtid() { return threadIdx.x < DS_WORK ? (int)threadIdx.x : (1 << 28); }

enum { DK_NORMAL = 0, DK_MASKED = 1, DK_ZERO = 2, DK_NULL = 3 };

// ------------------------------------------------------------------------------------------------
// shared-memory layout (same arithmetic on host and device)
// ------------------------------------------------------------------------------------------------
struct DsLayout {
  int streams, act, red, S, S2, pm, keys, ckeys, qs, wt, part, outs, lnp, bias, shs, nullkv, desc, kvbar, kvs, kv_rs3, kv_rsx,
      kvs_bytes, total;
};
__host__ __device__ inline int ds_align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline DsLayout ds_layout(int B, int D, int kmax, int H, int dh, int j3max, int nk) {
  DsLayout L;
  const int nt = (B + 7) / 8;
  const int jmax = j3max > nk + 1 ? j3max : nk + 1;
  int o = 0;
  L.streams = o; o += ds_align16(2 * B * D * 4);            // fp32 [2][B][D]
  L.act = o;     o += ds_align16(B * kmax * 2);             // bf16 [B][kmax] operand rows of the current product
  L.red = o;     o += ds_align16(DS_WARPS * 2 * nt * 128 * 4);  // partial 16 x 8 tiles of the K slices (value | gate)
  L.S = o;       o += ds_align16(H * jmax * 4);             // scores / probabilities [H][J]
  L.S2 = o;      o += ds_align16(H * j3max * 4);            // talking-heads output of the 3DNA window [H][J]
  L.pm = o;      o += ds_align16(jmax * 4);                 // mixed probabilities of one head
  L.keys = o;    o += ds_align16(j3max * 4);                // 3DNA key list of this sub-block: (kind << 28) | row
  L.ckeys = o;   o += ds_align16((nk + 1) * 4);             // cross-attention key list of this CTA's sample (whole launch)
  L.qs = o;      o += ds_align16(H * dh * 4);               // scaled query, fp32
  L.wt = o;      o += ds_align16(H * H * 4);                // talking-heads matrix
  L.part = o;    o += ds_align16(DS_THREADS * 2 * 4);       // PV partial sums
  L.outs = o;    o += ds_align16(H * dh * 4);               // attention output, fp32
  L.lnp = o;     o += ds_align16(4 * D * 4);                // post_w, post_b, pre_w, pre_b of the coming norms (cp.async)
  L.bias = o;    o += ds_align16(D * 4);                    // to_out bias of the current Sparse3DNA sub-block
  L.shs = o;     o += ds_align16(B * (D / 2) * 2);          // shifted channel quarters of the ShiftVideoTokens gather
  L.nullkv = o;  o += ds_align16(2 * dh * 4);               // learned null key / value of this CTA's head, fp32
  L.desc = o;    o += 4 * ds_align16((int)sizeof(nuwa_decode_sub));  // previous / current / next / next-but-one descriptor
  L.kvbar = o;   o += 32;                                   // mbarrier of the K / V bulk copies (8 B), slot of the new token in the 3DNA window (int), mbarrier of the norm-parameter bulk copies (8 B at +16)
  // staged K | V rows: 3DNA window of one sample (all heads) or the context head slices of one (sample, head);
  // row strides padded by 16 bytes so that lanes reading consecutive rows hit different banks
  L.kv_rs3 = H * dh * 2 + 16;
  L.kv_rsx = dh * 2;  // dense rows (one bulk copy per slice); readers rotate their 16-byte chunks by the row index
  const int kv3 = 2 * j3max * L.kv_rs3, kvx = nk > 0 ? 2 * nk * L.kv_rsx : 0;
  L.kvs_bytes = ds_align16(kv3 > kvx ? kv3 : kvx);
  L.kvs = o;     o += L.kvs_bytes;
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

// Measured alternatives (tools/decode_fence_ab.py, barriers-only mode, 2.9 us per barrier): relaxed polls + one acquire
// fence, __nanosleep between polls, red instead of atom -- all within 3 % of each other; the cost is the L2 round trips.
__device__ __forceinline__ void grid_barrier(GridBar& g) {
  __syncthreads();
  if (threadIdx.x == DS_WORK) {  // lane 0 of the sync warp
    g.target += g.nblocks;
    // release: everything this CTA wrote (ordered before by the CTA barrier) is visible to whoever acquires the count
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(g.count) : "memory");
    if (old + 1u == g.target) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");  // the last arriver has nothing to wait for
    } else {
      long long t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_u32(g.count) < g.target) {
        ++spins;
        if (spins == 4096u) t0 = clock64();
        // watchdog: a protocol bug becomes a launch error instead of a hung GPU (~2 s)
        if (spins > 4096u && (spins & 4095u) == 0u && (clock64() - t0) > 4000000000LL) __trap();
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no registers, one instruction per row), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most N of the most recently committed groups are still pending
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// skinny products on mma.sync:  out[n][tile*16 + r] = sum_k W[tile*16 + r][k] * As[n][k]
//   A operand = 16 weight rows, B operand = batch rows (N = 8 per MMA), fp32 accumulators.
//   k-block = 32 contraction indices; lane (g = lane/4, tg = lane%4) owns the 8 contiguous k [32*kb + 8*tg, +8) of weight
//   rows g and g+8 (one 16-byte load each) and of batch row g (one 16-byte shared-memory load): the two MMAs of a
//   k-block consume the .x/.y and .z/.w halves, i.e. the same permutation of k on both operands.
//   task = (tile, K-slice); the S slices of a tile sit in consecutive worker warps of one CTA (S = 1 or DS_WWARPS).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16x8x16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KB, bool PAIR>
struct WPre {
  uint4 a[KB][2];
  uint4 g[PAIR ? KB : 1][2];
};

struct GemvOut {
  float* f32;        // fp32 output or NULL
  bf16* b16;         // bf16 output or NULL
  long long ld;      // row stride (elements)
  const float* bias; // per output column (shared memory) or NULL
};

struct GemvTask {
  int tile, kb0, kb1;
  bool active;
};
__device__ __forceinline__ GemvTask gemv_task(int task, int ntiles, int K, int S) {
  GemvTask t;
  t.active = task < ntiles * S;
  t.tile = t.active ? task / S : 0;
  const int ks = t.active ? task - t.tile * S : 0;
  const int nkb = (K + 31) >> 5, per = (nkb + S - 1) / S;
  t.kb0 = ks * per;
  t.kb1 = min(nkb, t.kb0 + per);
  return t;
}

template <int KB, bool PAIR>
__device__ __forceinline__ void gemv_prefetch(WPre<KB, PAIR>& w, const bf16* __restrict__ W, int ldw, int ntiles, int K,
                                              int S, int gwarp, int lane) {
  const GemvTask t = gemv_task(gwarp, ntiles, K, S);
  const int g = lane >> 2, tg = lane & 3;
  const bf16* w0 = W + (long long)(t.tile * (PAIR ? 32 : 16) + g) * ldw + 8 * tg;
#pragma unroll
  for (int i = 0; i < KB; ++i) {
    const int kb = t.kb0 + i;
    const bool valid = t.active && kb < t.kb1 && kb * 32 + 8 * tg < K;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    w.a[i][0] = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + kb * 32)) : z;
    w.a[i][1] = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)8 * ldw + kb * 32)) : z;
    if (PAIR) {
      w.g[i][0] = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)16 * ldw + kb * 32)) : z;
      w.g[i][1] = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)24 * ldw + kb * 32)) : z;
    }
  }
}

__device__ __forceinline__ void gemv_emit(const GemvOut& o, int tile, int e, float v, float gt, bool pair, int B,
                                          int ncols) {
  const int nt = e >> 7, ln = (e & 127) >> 2, r = e & 3;
  const int row = (ln >> 2) + ((r & 2) ? 8 : 0);
  const int n = nt * 8 + 2 * (ln & 3) + (r & 1);
  const int col = tile * 16 + row;
  if (n >= B || col >= ncols) return;
  if (pair) v = v * gelu_erf(gt);
  else if (o.bias != nullptr) v += o.bias[col];
  if (o.f32 != nullptr) o.f32[(long long)n * o.ld + col] = v;
  if (o.b16 != nullptr) o.b16[(long long)n * o.ld + col] = __float2bfloat16(v);
}

// All threads of the CTA must call this (it contains __syncthreads when S > 1).  ncols = valid output columns.
template <int NT, int KB, bool PAIR>
__device__ __forceinline__ void gemv_run(const WPre<KB, PAIR>& wp, const bf16* __restrict__ W, int ldw, int ntiles,
                                         int ncols, int K, int S, const bf16* As, int lda_s, int B, float* red,
                                         const GemvOut& o, int gwarp, int total_warps, int warp, int lane) {
  const int rounds = (ntiles * S + total_warps - 1) / total_warps;
  const int g = lane >> 2, tg = lane & 3;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int r = 0; r < rounds; ++r) {
    const GemvTask t = gemv_task(gwarp + r * total_warps, ntiles, K, S);
    float c[NT][4], cg[PAIR ? NT : 1][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        c[n][k] = 0.f;
        if (PAIR) cg[n][k] = 0.f;
      }
    if (t.active) {
      const bf16* w0 = W + (long long)(t.tile * (PAIR ? 32 : 16) + g) * ldw + 8 * tg;
      auto step = [&](const uint4& a0, const uint4& a1, const uint4& g0, const uint4& g1, int kk) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int row = n * 8 + g;
          const uint4 av = (row < B && kk < K) ? *reinterpret_cast<const uint4*>(As + row * lda_s + kk) : z;
          mma_bf16_16x8x16(c[n], a0.x, a1.x, a0.y, a1.y, av.x, av.y);
          mma_bf16_16x8x16(c[n], a0.z, a1.z, a0.w, a1.w, av.z, av.w);
          if (PAIR) {
            mma_bf16_16x8x16(cg[n], g0.x, g1.x, g0.y, g1.y, av.x, av.y);
            mma_bf16_16x8x16(cg[n], g0.z, g1.z, g0.w, g1.w, av.z, av.w);
          }
        }
      };
      if (r == 0) {
#pragma unroll
        for (int i = 0; i < KB; ++i)
          if (t.kb0 + i < t.kb1) step(wp.a[i][0], wp.a[i][1], wp.g[PAIR ? i : 0][0], wp.g[PAIR ? i : 0][1], (t.kb0 + i) * 32 + 8 * tg);
      }
      for (int kb = t.kb0 + (r == 0 ? KB : 0); kb < t.kb1; ++kb) {  // beyond the prefetched blocks / later rounds
        const int kk = kb * 32 + 8 * tg;
        const bool valid = kk < K;
        const uint4 a0 = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + kb * 32)) : z;
        const uint4 a1 = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)8 * ldw + kb * 32)) : z;
        uint4 g0 = z, g1 = z;
        if (PAIR) {
          g0 = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)16 * ldw + kb * 32)) : z;
          g1 = valid ? __ldg(reinterpret_cast<const uint4*>(w0 + (long long)24 * ldw + kb * 32)) : z;
        }
        step(a0, a1, g0, g1, kk);
      }
    }
    if (S == 1) {
      if (t.active) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int k = 0; k < 4; ++k) gemv_emit(o, t.tile, n * 128 + lane * 4 + k, c[n][k], PAIR ? cg[n][k] : 0.f, PAIR, B, ncols);
      }
    } else {
      // partial tiles -> shared memory: red[warp][value|gate][NT*128]
      if (warp < DS_WWARPS) {
        float* mine = red + (size_t)warp * 2 * NT * 128;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          *reinterpret_cast<float4*>(mine + n * 128 + lane * 4) = make_float4(c[n][0], c[n][1], c[n][2], c[n][3]);
          if (PAIR) *reinterpret_cast<float4*>(mine + NT * 128 + n * 128 + lane * 4) = make_float4(cg[n][0], cg[n][1], cg[n][2], cg[n][3]);
        }
      }
      __syncthreads();
      const int tiles_here = DS_WWARPS / S;  // tiles of this CTA in this round
      const int first_task = blockIdx.x * DS_WWARPS + r * total_warps;
      for (int idx = threadIdx.x; idx < tiles_here * NT * 128; idx += DS_THREADS) {
        const int tl = idx / (NT * 128), e = idx - tl * NT * 128;
        const int task0 = first_task + tl * S;
        if (task0 < ntiles * S) {
          float v = 0.f, gt = 0.f;
          for (int s2 = 0; s2 < S; ++s2) {
            const float* src = red + (size_t)(tl * S + s2) * 2 * NT * 128;
            v += src[e];
            if (PAIR) gt += src[NT * 128 + e];
          }
          gemv_emit(o, task0 / S, e, v, gt, PAIR, B, ncols);
        }
      }
      __syncthreads();
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

// Requested well before the barrier that precedes the norms (bulk copies L2 -> shared memory issued by ONE thread, no
// registers held; as 16-byte cp.async pieces the request itself cost ~1.2 us of issue work per sub-block):
// lnp[0..1] = post-norm weight / bias of `prev`, lnp[2..3] = pre-norm weight / bias of `cur` (w2 / b2 when cur is
// NULL: the final StableLayerNorm), shs[b][0:D/2] = the two shifted channel quarters of ShiftVideoTokens, taken from the
// pre-norm rows of positions t - fmap and t - 1 (zeros at the grid border).  Completes one phase of `lnbar`.
__device__ __forceinline__ void prefetch_norms(const DecParams& p, const DecSub* prev, const DecSub* cur, const float* w2,
                                               const float* b2, int t, float* lnp, bf16* shs, uint64_t* lnbar) {
  const int D = p.D;
  const bool shift = cur != nullptr && cur->shift && t >= 1;
  int src_h = -1, src_w = -1;
  if (shift) {
    const int T = p.fmap * p.fmap;
    const int pos = (t - 1) % T;
    const int row = pos / p.fmap, col = pos - row * p.fmap;
    src_h = row > 0 ? t - p.fmap : -1;
    src_w = col > 0 ? t - 1 : -1;
  }
  const int q4 = D / 4;
  // one bulk copy per thread of warps 0-1 (a bulk-copy instruction costs its issuing thread ~150-250 cycles: issued by
  // one thread, the ~20 copies were 3 us on the critical path); every one of the 64 threads arrives on lnbar
  if (threadIdx.x < 64) {
    const int i = threadIdx.x;
    uint32_t tx = 0;
    if (i < 4) {
      const float* src = i == 0 ? (prev ? prev->post_w : nullptr)
                       : i == 1 ? (prev ? prev->post_b : nullptr)
                       : i == 2 ? (cur ? cur->pre_w : w2)
                                : (cur ? cur->pre_b : b2);
      if (src != nullptr) {
        fence_proxy_async_smem();
        bulk_g2s(lnp + i * D, src, (uint32_t)D * 4, lnbar);
        tx = (uint32_t)D * 4;
      }
    } else if (shift && i - 4 < 2 * p.B) {
      const int b = (i - 4) >> 1, quarter = (i - 4) & 1;
      const int src = quarter == 0 ? src_h : src_w;
      if (src >= 0) {
        const bf16* sc = reinterpret_cast<const bf16*>(cur->shift_cache);
        fence_proxy_async_smem();  // the zero-filled border quarters of an earlier sub-block were generic-proxy writes
        bulk_g2s(shs + b * (D / 2) + quarter * q4, sc + ((long long)b * p.npos + src) * D + quarter * q4, (uint32_t)q4 * 2, lnbar);
        tx = (uint32_t)q4 * 2;
      }
    }
    if (tx) mbar_arrive_expect_tx(lnbar, tx);
    else mbar_arrive(lnbar);
  }
  if (shift && (src_h < 0 || src_w < 0)) {  // grid border: the shifted-in quarter is zero
    const int qp = D / 32;  // 16-byte pieces per channel quarter (bf16)
    for (int i = ds_wtid(); i < p.B * 2 * qp; i += DS_WORK) {
      const int b = i / (2 * qp), r = i - b * 2 * qp;
      const int quarter = r / qp, k = r - quarter * qp;
      if ((quarter == 0 ? src_h : src_w) < 0)
        *reinterpret_cast<uint4*>(shs + b * (D / 2) + quarter * q4 + k * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// prev != NULL:  streams[prev->write] += LayerNorm_post(y)        (SandwichNorm tail + residual)
// cur  != NULL:  As = bf16(LayerNorm_pre(streams[cur->read])) with the ShiftVideoTokens gather
// Waits for the phase `ln_parity` of `lnbar` (the bulk copies of prefetch_norms).
__device__ __forceinline__ void ln_prologue(const DecParams& p, const DecSub* prev, const DecSub* cur, int t, float* streams,
                                            bf16* As, int lda_s, const float* lnp, const bf16* shs, int warp, int lane,
                                            uint64_t* lnbar, uint32_t ln_parity) {
  const int D = p.D, B = p.B;
  float4 v[DS_LNV];
  int nv = 0;
#pragma unroll
  for (int i = 0; i < DS_LNV; ++i)
    if ((lane + 32 * i) * 4 < D) nv = i + 1;
  auto load_y = [&](int b) {
#pragma unroll
    for (int i = 0; i < DS_LNV; ++i)
      if (i < nv) v[i] = __ldcg(reinterpret_cast<const float4*>(p.y + (long long)b * D + (lane + 32 * i) * 4));
  };
  if (warp < B && prev != nullptr) load_y(warp);  // the only loads that had to wait for the barrier
  mbar_wait(lnbar, ln_parity);
  __syncthreads();
  for (int b = warp; b < B; b += DS_WARPS) {
    if (prev != nullptr) {
      if (b != warp) load_y(b);
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
  for (int b = warp; b < B; b += DS_WARPS) {
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
// attention pieces (one query row per sample); K / V rows are read from shared memory
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

// out[c] = P[0] * null_v[c] + sum_{j=1..n} P[j] * V_{j-1}[c] for one head (dh channels): every thread takes a channel pair
// and every KG-th key, no branches (masked slots carry P = 0 and finite stale V), partial sums reduced through `part`
__device__ __forceinline__ void attn_pv_head(const float* P, const uint8_t* Vrows, int rs, const float* null_v, int dh, int n,
                                             float* part, float* outs) {
  const int npairs = dh / 2;
  const int KG = DS_THREADS / npairs;
  const int cp = threadIdx.x % npairs, kg = threadIdx.x / npairs;
  if (kg < KG) {
    const uint8_t* vcol = Vrows + cp * 4;
    float ax = 0.f, ay = 0.f;
#pragma unroll 8
    for (int j = kg; j < n; j += KG) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vcol + (size_t)j * rs));
      const float pj = P[j + 1];
      ax = fmaf(pj, v.x, ax);
      ay = fmaf(pj, v.y, ay);
    }
    part[(kg * npairs + cp) * 2 + 0] = ax;
    part[(kg * npairs + cp) * 2 + 1] = ay;
  }
  __syncthreads();
  if (threadIdx.x < npairs) {
    float ax = P[0] * null_v[2 * threadIdx.x], ay = P[0] * null_v[2 * threadIdx.x + 1];
    for (int g = 0; g < KG; ++g) {
      ax += part[(g * npairs + threadIdx.x) * 2 + 0];
      ay += part[(g * npairs + threadIdx.x) * 2 + 1];
    }
    outs[2 * threadIdx.x] = ax;
    outs[2 * threadIdx.x + 1] = ay;
  }
  __syncthreads();
}

// K / V prefetch of a sub-block's attention (consumed two barriers later).
//   3DNA:  CTA b < B stages the window rows of sample b that earlier steps wrote (row t, the new token, comes later):
//          ONE bulk copy per K row and per V row (all heads: inner * 2 contiguous bytes), issued by the thread that owns
//          the key slot -- a per-16-byte cp.async loop here was ~5 us of issue work on the critical path of every CTA
//          that waits for these B CTAs at the next barrier
//   cross: CTA w < B*H stages the context K and V slices of its (sample, head) (dh * 2 bytes per row), the head's
//          null key / value
//   both:  the talking-heads matrix; 3DNA (every CTA): the to_out bias (cp.async)
// The bulk copies complete on the `kvbar` mbarrier (every worker thread arrives once, with the bytes it requested);
// slots that are not loaded (masked / zero keys) keep stale but finite bf16 data: their probabilities are forced to
// zero before PV and their scores are replaced, so the contents never matter.  Commits exactly one cp.async group.
__device__ __forceinline__ void prefetch_kv(const DecParams& p, const DecSub* s, const DsLayout& L, uint8_t* smem, int t) {
  const int H = p.H, dh = p.dh, inner = H * dh;
  uint64_t* kvbar = reinterpret_cast<uint64_t*>(smem + L.kvbar);
  if (s->kind == NUWA_DEC_3DNA) {
    if (s->b_out != nullptr)  // every CTA adds the to_out bias in phase 3
      for (int i = ds_wtid(); i < p.D / 4; i += DS_WORK) cp_async16(smem + L.bias + i * 16, s->b_out + i * 4);
    if ((int)blockIdx.x < p.B && t > 0) {
      const int b = blockIdx.x;
      int* keys = reinterpret_cast<int*>(smem + L.keys);
      int* self_slot = reinterpret_cast<int*>(smem + L.kvbar + 8);
      const int J = 1 + s->kt * s->kh * s->kw;
      if (threadIdx.x == 0) *self_slot = -1;
      __syncthreads();
      for (int j = threadIdx.x; j < J; j += DS_THREADS) {
        int row = 0;
        const int kd = key3dna(p, *s, t, t, j, row);
        keys[j] = (kd << 28) | row;
        if (kd == DK_NORMAL && row == t) *self_slot = j;  // the new token's own slot (at most one: offsets are distinct)
      }
      __syncthreads();
      const bf16* cb = reinterpret_cast<const bf16*>(s->cache) + (long long)b * p.npos * 3 * inner;
      uint8_t* Ks = smem + L.kvs;
      uint8_t* Vs = Ks + (size_t)J * L.kv_rs3;
      if (threadIdx.x < DS_WORK) {
        fence_proxy_async_smem();  // earlier generic-proxy reads / writes of the staging area come before the bulk writes
        uint32_t tx = 0;
        for (int j = threadIdx.x; j < J; j += DS_WORK) {
          const int kj = keys[j];
          const int row = kj & 0x0FFFFFFF;
          if ((kj >> 28) == DK_NORMAL && row != t) {
            const bf16* src = cb + (long long)row * 3 * inner + inner;
            bulk_g2s(Ks + (size_t)j * L.kv_rs3, src, (uint32_t)inner * 2, kvbar);
            bulk_g2s(Vs + (size_t)j * L.kv_rs3, src + inner, (uint32_t)inner * 2, kvbar);
            tx += (uint32_t)inner * 4;
          }
        }
        if (tx) mbar_arrive_expect_tx(kvbar, tx);
        else mbar_arrive(kvbar);
      }
      for (int i = ds_wtid(); i < H * H / 4; i += DS_WORK)
        cp_async16(smem + L.wt + i * 16, s->talk + i * 4);
    }
  } else if (s->kind == NUWA_DEC_CROSS) {
    if ((int)blockIdx.x < p.B * H) {
      const int b = blockIdx.x / H, h = blockIdx.x - b * H;
      // context K / V of this (sample, head): packed [B][H][2][nk][dh] by the host (engine.FusedDecode), so each slice is
      // ONE contiguous bulk copy; masked rows come along (their probabilities are exactly 0)
      const bf16* kv = reinterpret_cast<const bf16*>(s->cache) + ((long long)b * H + h) * 2 * p.nk * dh;
      uint8_t* Ks = smem + L.kvs;
      uint8_t* Vs = Ks + (size_t)p.nk * L.kv_rsx;
      if (threadIdx.x < DS_WORK) {
        const uint32_t bytes = (uint32_t)p.nk * dh * 2;
        if (threadIdx.x == 0) {
          fence_proxy_async_smem();  // generic-proxy accesses of the staging area (made visible by the CTA barriers) first
          bulk_g2s(Ks, kv, bytes, kvbar);
          bulk_g2s(Vs, kv + (long long)p.nk * dh, bytes, kvbar);
          mbar_arrive_expect_tx(kvbar, 2 * bytes);
        } else {
          mbar_arrive(kvbar);
        }
      }
      float* nkv = reinterpret_cast<float*>(smem + L.nullkv);
      for (int i = ds_wtid(); i < 2 * (dh / 4); i += DS_WORK) {
        const int sel = i / (dh / 4), k = i - sel * (dh / 4);
        cp_async16(nkv + sel * dh + k * 4, (sel ? s->null_v : s->null_k) + h * dh + k * 4);
      }
      for (int i = ds_wtid(); i < H * H / 4; i += DS_WORK)
        cp_async16(smem + L.wt + i * 16, s->talk + i * 4);
    }
  }
  cp_async_commit();
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
#define DS_STAMP(id)                                                            \
  do {                                                                          \
    if (p.prof != nullptr && blockIdx.x == p.prof_cta && threadIdx.x == 0) {   \
      p.prof[2 * nstamp] = (long long)(si * 32 + (id));                         \
      p.prof[2 * nstamp + 1] = clock64();                                       \
      ++nstamp;                                                                 \
    }                                                                           \
  } while (0)

template <int NT>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_stack_kernel(const DecParams p) {
  extern __shared__ __align__(16) uint8_t ds_smem[];
  const int B = p.B, D = p.D, H = p.H, dh = p.dh, inner = H * dh;
  const DsLayout L = ds_layout(B, D, p.kmax, H, dh, p.j3max, p.nk);
  float* streams = reinterpret_cast<float*>(ds_smem + L.streams);
  bf16* As = reinterpret_cast<bf16*>(ds_smem + L.act);
  float* red = reinterpret_cast<float*>(ds_smem + L.red);
  float* Ss = reinterpret_cast<float*>(ds_smem + L.S);
  float* S2 = reinterpret_cast<float*>(ds_smem + L.S2);
  float* Pm = reinterpret_cast<float*>(ds_smem + L.pm);
  int* keys = reinterpret_cast<int*>(ds_smem + L.keys);
  int* ckeys = reinterpret_cast<int*>(ds_smem + L.ckeys);
  float* qs = reinterpret_cast<float*>(ds_smem + L.qs);
  float* Wt = reinterpret_cast<float*>(ds_smem + L.wt);
  float* part = reinterpret_cast<float*>(ds_smem + L.part);
  float* outs = reinterpret_cast<float*>(ds_smem + L.outs);
  float* lnp = reinterpret_cast<float*>(ds_smem + L.lnp);
  bf16* shs = reinterpret_cast<bf16*>(ds_smem + L.shs);
  float* nullkv = reinterpret_cast<float*>(ds_smem + L.nullkv);
  uint8_t* kvs = ds_smem + L.kvs;
  uint64_t* kvbar = reinterpret_cast<uint64_t*>(ds_smem + L.kvbar);
  const int* self_slot = reinterpret_cast<const int*>(ds_smem + L.kvbar + 8);
  uint64_t* lnbar = reinterpret_cast<uint64_t*>(ds_smem + L.kvbar + 16);
  uint32_t kv_uses = 0, ln_uses = 0;  // phases of kvbar / lnbar consumed by this CTA
  constexpr int DESC_STRIDE = (sizeof(DecSub) + 15) & ~15;
  constexpr int DESC_WORDS = sizeof(DecSub) / 4;
  const int lda_s = p.kmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = warp < DS_WWARPS ? blockIdx.x * DS_WWARPS + warp : (1 << 28);  // the sync warp takes no tasks
  const int total_warps = gridDim.x * DS_WWARPS;
  const int t = __ldg(p.t_ptr);
  const float qscale = rsqrtf((float)dh);
  const int dbg = p.debug_flags;
  const int S16 = p.split_small;  // K slices of the D- / inner-wide products
  GridBar bar{p.barrier, gridDim.x, 0u};
  bf16* act = reinterpret_cast<bf16*>(p.act);
  bf16* actq = reinterpret_cast<bf16*>(p.actq);
  int nstamp = 0;
  int si = 0;

  // X = x (plain) or [x, x] (reversible.py:133); descriptors of the first two sub-blocks; cross-attention key list
  for (int i = threadIdx.x; i < B * D; i += DS_THREADS) {
    const float v = __ldg(p.x_in + i);
    streams[i] = v;
    if (p.reversible) streams[B * D + i] = v;
  }
  if (threadIdx.x < DESC_WORDS) {
    reinterpret_cast<uint32_t*>(ds_smem + L.desc)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(p.subs) + threadIdx.x);
    if (p.nsubs > 1)
      reinterpret_cast<uint32_t*>(ds_smem + L.desc + DESC_STRIDE)[threadIdx.x] =
          __ldg(reinterpret_cast<const uint32_t*>(p.subs + 1) + threadIdx.x);
  }
  // the K / V staging area only ever holds bf16 data written by the bulk copies; start it from zeros so that slots
  // that are never loaded (masked keys) read as finite values
  for (int i = threadIdx.x; i < L.kvs_bytes / 16; i += DS_THREADS) reinterpret_cast<uint4*>(kvs)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(kvbar, DS_WORK);
    mbar_init(lnbar, 64);
    fence_barrier_init();
  }
  if (p.nk > 0 && (int)blockIdx.x < B * H) {
    const int b = blockIdx.x / H;
    for (int j = threadIdx.x; j <= p.nk; j += DS_THREADS) {
      int kd = DK_NULL, row = 0;
      if (j > 0) {
        row = j - 1;
        kd = (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + row] == 0) ? DK_MASKED : DK_NORMAL;
      }
      ckeys[j] = (kd << 28) | row;
    }
  }
  __syncthreads();
  {
    const DecSub* s0 = reinterpret_cast<const DecSub*>(ds_smem + L.desc);
    prefetch_norms(p, nullptr, s0, nullptr, nullptr, t, lnp, shs, lnbar);
  }
  DS_STAMP(0);

  const DecSub* prev = nullptr;
  for (si = 0; si < p.nsubs; ++si) {
    const DecSub* s = reinterpret_cast<const DecSub*>(ds_smem + L.desc + (si % 4) * DESC_STRIDE);
    const DecSub* nxt = si + 1 < p.nsubs ? reinterpret_cast<const DecSub*>(ds_smem + L.desc + ((si + 1) % 4) * DESC_STRIDE) : nullptr;
    // descriptor of sub-block si + 2: requested now, parked in shared memory after the first barrier wait
    uint32_t next_word = 0;
    const bool has_next2 = si + 2 < p.nsubs && threadIdx.x < DESC_WORDS;  // worker warps 0-1
    if (has_next2) next_word = __ldg(reinterpret_cast<const uint32_t*>(p.subs + si + 2) + threadIdx.x);
    const int kind = s->kind;
    const bf16* Wa = reinterpret_cast<const bf16*>(s->w_a);
    const bf16* Wb = reinterpret_cast<const bf16*>(s->w_b);
    // one cp.async group per sub-block (small attention parameters, committed by prefetch_kv); K / V rows and the norm
    // parameters travel as bulk copies on their own mbarriers
    auto phase1 = [&]() {
      prefetch_kv(p, s, L, ds_smem, t);
      DS_STAMP(1);
      if (prev != nullptr) grid_barrier(bar);  // y of the previous sub-block is complete
      DS_STAMP(2);
      if (has_next2)
        reinterpret_cast<uint32_t*>(ds_smem + L.desc + ((si + 2) % 4) * DESC_STRIDE)[threadIdx.x] = next_word;
      if (!(dbg & 1)) ln_prologue(p, prev, s, t, streams, As, lda_s, lnp, shs, warp, lane, lnbar, ln_uses & 1u);
      else { mbar_wait(lnbar, ln_uses & 1u); __syncthreads(); }
      ++ln_uses;
      DS_STAMP(10);
      // norm parameters of the NEXT sub-block (or of the final StableLayerNorm): >= two barriers ahead of their use
      prefetch_norms(p, s, nxt, p.norm_w, p.norm_b, t, lnp, shs, lnbar);
      DS_STAMP(3);
    };
    if (kind == NUWA_DEC_3DNA) {
      WPre<1, false> wa;
      gemv_prefetch(wa, Wa, D, 3 * inner / 16, D, S16, gwarp, lane);
      phase1();
      // q|k|v of the new token go straight into row t of the cache
      bf16* cache = reinterpret_cast<bf16*>(s->cache);
      GemvOut o{nullptr, cache + (long long)t * 3 * inner, (long long)p.npos * 3 * inner, nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wa, Wa, D, 3 * inner / 16, 3 * inner, D, S16, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      WPre<1, false> wb;
      gemv_prefetch(wb, Wb, inner, D / 16, inner, S16, gwarp, lane);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      // ---- phase 2: attention, one CTA per sample (all heads: talking heads mix across heads) ----
      if (!(dbg & 4) && (int)blockIdx.x < B) {
        const int b = blockIdx.x;
        const bf16* cb = cache + (long long)b * p.npos * 3 * inner;
        if (t == 0) {  // bos attends only to itself (:499,608)
          for (int c = threadIdx.x; c < inner; c += DS_THREADS) act[(long long)b * inner + c] = __ldcg(cb + 2 * inner + c);
        } else {
          const int J = 1 + s->kt * s->kh * s->kw;
          uint8_t* Ks = kvs;
          uint8_t* Vs = kvs + (size_t)J * L.kv_rs3;
          // the new token's own q (scaled, fp32) and its k / v into their window slot: one L2 round trip
          {
            const int self = *self_slot;
            const int pieces = inner / 8;
            const bf16* rowt = cb + (long long)t * 3 * inner;
            for (int i = threadIdx.x; i < 3 * pieces; i += DS_THREADS) {
              const int sel = i / pieces, k = i - sel * pieces;
              if (sel == 0) {
                const uint4 u = __ldcg(reinterpret_cast<const uint4*>(rowt + k * 8));
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                *reinterpret_cast<float4*>(qs + k * 8) = make_float4(f0.x * qscale, f0.y * qscale, f1.x * qscale, f1.y * qscale);
                *reinterpret_cast<float4*>(qs + k * 8 + 4) = make_float4(f2.x * qscale, f2.y * qscale, f3.x * qscale, f3.y * qscale);
              } else if (self >= 0) {
                *reinterpret_cast<uint4*>((sel == 2 ? Vs : Ks) + (size_t)self * L.kv_rs3 + k * 16) =
                    __ldcg(reinterpret_cast<const uint4*>(rowt + sel * inner + k * 8));
              }
            }
          }
          DS_STAMP(16);
          mbar_wait(kvbar, kv_uses & 1u);   // window rows of the earlier tokens (bulk copies)
          cp_async_wait<0>();               // talking-heads matrix, bias
          __syncthreads();
          DS_STAMP(17);
          // scores + softmax: warp = head, lane = key slot (no CTA barrier in between; loads are unconditional, the key
          // kind only selects the value)
          for (int h = warp; h < H; h += DS_WARPS) {
            const float* q = qs + h * dh;
            float* Sh = Ss + h * J;
            float m = -FLT_MAX;
            for (int j = lane; j < J; j += 32) {
              const int kind = keys[j] >> 28;
              const uint4* kr = reinterpret_cast<const uint4*>(Ks + (size_t)j * L.kv_rs3 + h * dh * 2);
              float sc = 0.f;
              for (int i = 0; i < dh / 8; ++i) sc += dot8q(q + i * 8, kr[i]);
              sc = kind == DK_NORMAL ? sc : (kind == DK_MASKED ? -FLT_MAX : 0.f);
              Sh[j] = sc;
              m = fmaxf(m, sc);
            }
            m = warp_max(m);
            float l = 0.f;
            for (int j = lane; j < J; j += 32) {
              const float e = __expf(Sh[j] - m);
              Sh[j] = e;
              l += e;
            }
            l = warp_sum(l);
            const float inv = 1.0f / l;
            for (int j = lane; j < J; j += 32) Sh[j] *= inv;
          }
          __syncthreads();
          DS_STAMP(19);
          // talking heads (:556-558), one output per thread; slots without a real key are dropped from PV here
          // (masked: P = 0 already; zero keys: V = 0)
          for (int item = threadIdx.x; item < H * J; item += DS_THREADS) {
            const int gh = item / J, j = item - gh * J;
            float a = 0.f;
#pragma unroll 8
            for (int h = 0; h < H; ++h) a = fmaf(Wt[gh * H + h], Ss[h * J + j], a);
            S2[item] = (keys[j] >> 28) == DK_NORMAL ? a : 0.f;
          }
          __syncthreads();
          DS_STAMP(20);
          // PV: thread = channel pair, branch-free over the window
          for (int cp = threadIdx.x; cp < inner / 2; cp += DS_THREADS) {
            const int hl = cp / (dh / 2);
            const float* Ph = S2 + hl * J;
            const uint8_t* vcol = Vs + cp * 4;
            float ax = 0.f, ay = 0.f;
#pragma unroll 8
            for (int j = 0; j < J; ++j) {
              const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vcol + (size_t)j * L.kv_rs3));
              ax = fmaf(Ph[j], v.x, ax);
              ay = fmaf(Ph[j], v.y, ay);
            }
            *reinterpret_cast<uint32_t*>(act + (long long)b * inner + 2 * cp) = pack_bf16x2(ax, ay);
          }
          DS_STAMP(21);
        }
        if (t > 0) ++kv_uses;
      }
      DS_STAMP(6);
      grid_barrier(bar);
      DS_STAMP(7);
      // ---- phase 3: to_out (+ bias) ----
      cp_async_wait<0>();  // the bias (kv group); made visible by the CTA barrier inside stage_act
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, s->b_out != nullptr ? reinterpret_cast<const float*>(ds_smem + L.bias) : nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wb, Wb, inner, D / 16, D, inner, S16, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    } else if (kind == NUWA_DEC_CROSS) {
      WPre<1, false> wa;
      gemv_prefetch(wa, Wa, D, inner / 16, D, S16, gwarp, lane);
      phase1();
      GemvOut o{nullptr, actq, (long long)inner, nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wa, Wa, D, inner / 16, inner, D, S16, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      WPre<1, false> wb;
      gemv_prefetch(wb, Wb, inner, D / 16, inner, S16, gwarp, lane);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      const int J = p.nk + 1;
      // slot j >= 1 of the key list is context row j - 1, staged at row j - 1 of Ks / Vs: shift the bases by one row
      const uint8_t* Ks = kvs - L.kv_rsx;
      const bool mine = (int)blockIdx.x < B * H;
      const int b = blockIdx.x / H, h = blockIdx.x - b * H;
      // ---- phase 2a: probabilities of one (sample, head) per CTA (the softmax of a head needs only that head) ----
      if (!(dbg & 4) && mine) {
        for (int c = threadIdx.x; c < dh; c += DS_THREADS)
          qs[c] = __bfloat162float(__ldcg(actq + (long long)b * inner + h * dh + c)) * qscale;
        mbar_wait(kvbar, kv_uses & 1u);   // context K / V slices (bulk copies)
        cp_async_wait<0>();               // null key / value, talking-heads matrix
        __syncthreads();
        DS_STAMP(22);
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {  // thread = key slot; unconditional loads, the kind selects
          const int kind = ckeys[j] >> 28;
          float sc = 0.f;
          if (j == 0) {
            for (int i = 0; i < dh; ++i) sc = fmaf(qs[i], nullkv[i], sc);
          } else {
            const uint4* kr = reinterpret_cast<const uint4*>(Ks + (size_t)j * L.kv_rsx);
            const int nch = dh / 8;
            for (int i = 0; i < nch; ++i) {  // dense 128-byte rows: start at chunk j so that a quarter-warp spans all banks
              const int c = (i + j) % nch;
              sc += dot8q(qs + c * 8, kr[c]);
            }
          }
          Ss[j] = kind == DK_MASKED ? -FLT_MAX : sc;
        }
        __syncthreads();
        DS_STAMP(23);
        {  // every warp reduces the whole row (redundant, no further CTA barrier), every thread normalises its own slots
          float m = -FLT_MAX;
          for (int j = lane; j < J; j += 32) m = fmaxf(m, Ss[j]);
          m = warp_max(m);
          float l = 0.f;
          for (int j = lane; j < J; j += 32) l += __expf(Ss[j] - m);
          l = warp_sum(l);
          const float inv = 1.0f / l;
          for (int j = threadIdx.x; j < J; j += DS_THREADS) p.scores[((long long)b * H + h) * J + j] = __expf(Ss[j] - m) * inv;
        }
      }
      if (mine) ++kv_uses;
      DS_STAMP(6);
      grid_barrier(bar);
      DS_STAMP(7);
      // ---- phase 2b: talking-heads row h over the sample's probabilities, PV of head h ----
      if (!(dbg & 4) && mine) {
        for (int i = threadIdx.x; i < H * J; i += DS_THREADS) Ss[i] = __ldcg(p.scores + (long long)b * H * J + i);
        __syncthreads();
        DS_STAMP(24);
        for (int j = threadIdx.x; j < J; j += DS_THREADS) {  // :372 (masked slots: every head's probability is exactly 0)
          float a = 0.f;
#pragma unroll 8
          for (int hh = 0; hh < H; ++hh) a = fmaf(Wt[h * H + hh], Ss[hh * J + j], a);
          Pm[j] = a;
        }
        __syncthreads();
        DS_STAMP(26);
        attn_pv_head(Pm, kvs + (size_t)p.nk * L.kv_rsx, L.kv_rsx, nullkv + dh, dh, p.nk, part, outs);
        DS_STAMP(27);
        for (int c = threadIdx.x; c < dh; c += DS_THREADS) act[(long long)b * inner + h * dh + c] = __float2bfloat16(outs[c]);
      }
      DS_STAMP(8);
      grid_barrier(bar);
      DS_STAMP(9);
      stage_act(As, lda_s, act, inner, B, inner);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wb, Wb, inner, D / 16, D, inner, S16, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    } else {
      // GEGLU product: value / gate tile pairs -> a * gelu(g) (:255-258)
      const int ip = s->ip;
      WPre<1, true> wa;
      gemv_prefetch(wa, Wa, D, ip / 16, D, S16, gwarp, lane);
      phase1();
      GemvOut o{nullptr, act, (long long)ip, nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wa, Wa, D, ip / 16, ip, D, S16, As, lda_s, B, red, o, gwarp, total_warps, warp, lane);
      WPre<7, false> wb;
      gemv_prefetch(wb, Wb, ip, D / 16, ip, p.split_ff, gwarp, lane);
      DS_STAMP(4);
      grid_barrier(bar);
      DS_STAMP(5);
      stage_act(As, lda_s, act, ip, B, ip);
      GemvOut o3{p.y, nullptr, (long long)D, nullptr};
      if (!(dbg & 2)) gemv_run<NT>(wb, Wb, ip, D / 16, D, ip, p.split_ff, As, lda_s, B, red, o3, gwarp, total_warps, warp, lane);
      DS_STAMP(12);
    }
    prev = s;
  }
  // ---- tail: last post-norm + residual, StableLayerNorm, logits ----
  WPre<3, false> wl;
  const bf16* Wl = reinterpret_cast<const bf16*>(p.w_logits);
  if (Wl != nullptr) gemv_prefetch(wl, Wl, D, (p.V + 15) / 16, D, p.split_logits, gwarp, lane);
  grid_barrier(bar);
  DS_STAMP(13);
  ln_prologue(p, prev, nullptr, t, streams, As, lda_s, lnp, shs, warp, lane, lnbar, ln_uses & 1u);
  stable_ln_rows(p, streams, As, lda_s, lnp, warp, lane);
  if (Wl != nullptr) {
    GemvOut ol{p.logits, nullptr, (long long)p.V, nullptr};
    gemv_run<NT>(wl, Wl, D, (p.V + 15) / 16, p.V, D, p.split_logits, As, lda_s, B, red, ol, gwarp, total_warps, warp, lane);
  }
  DS_STAMP(14);
  cp_async_wait<0>();
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
static bool valid_split(int s) { return s == 1 || s == DS_WWARPS; }

int decode_stack(const DecParams& p_in, int cooperative, cudaStream_t stream) {
  DecParams p = p_in;
  if (p.subs == nullptr || p.nsubs <= 0 || p.B <= 0 || p.B > 16 || p.t_ptr == nullptr || p.barrier == nullptr)
    return NUWA_ERR_INVALID;
  if (p.D % 32 != 0 || p.D > 1024 || p.H <= 0 || p.H > 16 || p.dh % 8 != 0 || p.H * p.dh > 1024 || (p.H * p.dh) % 16 != 0 ||
      (p.H * p.H) % 4 != 0)
    return NUWA_ERR_INVALID;
  if (p.kmax < p.D || p.kmax < p.H * p.dh || p.kmax % 8 != 0) return NUWA_ERR_INVALID;
  if (p.x_in == nullptr || p.y == nullptr || p.act == nullptr || p.actq == nullptr || p.norm_w == nullptr ||
      p.norm_b == nullptr)
    return NUWA_ERR_INVALID;
  if (p.w_logits != nullptr && (p.logits == nullptr || p.V <= 0)) return NUWA_ERR_INVALID;
  if (p.j3max < 1 || p.j3max > 4096 || p.nk < 0) return NUWA_ERR_INVALID;
  p.jmax = p.j3max > p.nk + 1 ? p.j3max : p.nk + 1;
  if (p.nk > 0 && p.scores == nullptr) return NUWA_ERR_INVALID;
  if (!valid_split(p.split_small)) p.split_small = DS_WWARPS;
  if (!valid_split(p.split_ff)) p.split_ff = DS_WWARPS;
  if (!valid_split(p.split_logits)) p.split_logits = DS_WWARPS;
  const DsLayout L = ds_layout(p.B, p.D, p.kmax, p.H, p.dh, p.j3max, p.nk);
  if (L.total > DS_SMEM_MAX) return NUWA_ERR_INVALID;
  const void* kernel = p.B <= 8 ? (const void*)decode_stack_kernel<1> : (const void*)decode_stack_kernel<2>;
  static int max_blocks_per_sm_smem = -1, cached_smem = -1;
  static const void* cached_kernel = nullptr;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess) return NUWA_ERR_CUDA;
  if (cached_smem != L.total || cached_kernel != kernel) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, DS_THREADS, (size_t)L.total) != cudaSuccess)
      return NUWA_ERR_CUDA;
    max_blocks_per_sm_smem = nb;
    cached_smem = L.total;
    cached_kernel = kernel;
  }
  if (max_blocks_per_sm_smem < 1) return NUWA_ERR_INVALID;
  int grid = device_sm_count();
  if (p.max_ctas > 0 && p.max_ctas < grid) grid = p.max_ctas;
  // every sample (3DNA) / (sample, head) pair (cross attention) needs its own CTA: the K/V staging is per CTA
  const int need = p.nk > 0 ? p.B * p.H : p.B;
  if (grid < need) return NUWA_ERR_INVALID;
  if (cooperative) {
    void* args[] = {(void*)&p};
    if (cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(DS_THREADS), args, (size_t)L.total, stream) != cudaSuccess) {
      cudaGetLastError();
      return NUWA_ERR_CUDA;
    }
    ++g_launch_count;
    return NUWA_OK;
  }
  // plain launch: grid <= SM count and one CTA per SM, so all CTAs become resident once earlier work drains
  if (p.B <= 8) decode_stack_kernel<1><<<grid, DS_THREADS, L.total, stream>>>(p);
  else decode_stack_kernel<2><<<grid, DS_THREADS, L.total, stream>>>(p);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
