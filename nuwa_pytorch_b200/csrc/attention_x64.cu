// Dense attention core (Attention.forward, nuwa_pytorch.py:339-378) for the decoder's text cross-attention and the text
// encoder's self-attention at the model's native geometry: 8 heads x 64, learned null key, key mask, talking heads.
//
// The talking-heads 1x1 conv (:372) mixes the NORMALISED probabilities of all heads, so a flash-style single pass is not
// possible; attention_mma.cu therefore keeps a whole 16-query x 8-head probability slab on chip and re-reads K / V from L2
// for every 16 queries (673 MB of L2 traffic and 0.42 ms per launch at cfg 3).  This kernel makes the query tile 64 rows
// and runs two passes over the keys instead:
//
//   pass 1  row statistics (max, 1/sum) of every head: online softmax over 32-key chunks, nothing stored
//   pass 2  recompute the logits chunk by chunk, normalise, mix the heads IN REGISTERS (a lane of an m16n8 accumulator
//           holds the same (query, key) slots for every head), feed the mixed probabilities straight back into the
//           tensor cores as the A operand of P'V.
//
// CTA = 64 queries of one sample, 16 warps: warp = (16-query tile, pair of heads).  A warp computes the logits of its two
// heads, publishes the normalised probabilities (bf16) to the three other warps of its tile through shared memory (named
// barrier per tile), mixes all 8 heads for its two OUTPUT heads and accumulates their P'V (64 accumulator registers, so
// 16 warps fit an SM and hide the ldmatrix / mma latencies).  Q (once) and 16-key K / V chunks (double buffered cp.async)
// are staged in shared memory and read with ldmatrix; the learned null key stays an exact fp32 side column.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

namespace {

constexpr int XQ = 64;        // queries per CTA
constexpr int XK = 16;        // keys per staged chunk (one m16n8k16 contraction step of P'V)
constexpr int XH = 8, XD = 64, XC = XH * XD;
constexpr int XP = XC + 8;    // bf16 row pitch (1040 B): conflict-free ldmatrix rows
constexpr int XS = 3;         // cp.async stages of the K / V chunk ring
constexpr int XT = 512;       // threads: 16 warps = 4 query tiles x 4 head pairs
constexpr int EXP = 16 + 2;   // bf16 pitch of one exchanged probability row (16 keys), padded against bank conflicts

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// rows [r0, r0+nrows) of a (row-strided) bf16 matrix with XC columns -> dst[row][XP]; rows >= limit are zero filled
__device__ __forceinline__ void stage_rows(bf16* dst, const bf16* src, long long row_stride, int r0, int nrows, int limit) {
  for (int i = threadIdx.x; i < nrows * (XC / 8); i += blockDim.x) {
    const int r = i / (XC / 8), c = (i - r * (XC / 8)) * 8;
    const bool ok = (r0 + r) < limit;
    cp_async16(dst + r * XP + c, src + (ok ? (long long)(r0 + r) * row_stride + c : 0), ok);
  }
}

// logits of heads h0, h0+1 of this warp's 16 queries against the 16 staged keys: s[head][8-key block][c-fragment]
__device__ __forceinline__ void scores2(const bf16* Qs, const bf16* Kc, int rbase, int h0, int l8, int id, float (&s)[2][2][4]) {
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) s[hh][blk][0] = s[hh][blk][1] = s[hh][blk][2] = s[hh][blk][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qa[4], kb[4];
      ldsm4(qa, Qs + (rbase + l8 + 8 * (id & 1)) * XP + (h0 + hh) * XD + ks * 16 + 8 * (id >> 1));
      ldsm4(kb, Kc + (l8 + 8 * (id >> 1)) * XP + (h0 + hh) * XD + ks * 16 + 8 * (id & 1));
      mma16816(s[hh][0], qa, kb[0], kb[1]);
      mma16816(s[hh][1], qa, kb[2], kb[3]);
    }
  }
}

__global__ void __launch_bounds__(XT, 1) attn_dense_x64_kernel(const AttnParams p, int nk) {
  extern __shared__ __align__(16) uint8_t smem_x[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_x);                 // [XQ][XP]
  bf16* Ks = Qs + XQ * XP;                                    // [XS][XK][XP]
  bf16* Vs = Ks + XS * XK * XP;                               // [XS][XK][XP]
  bf16* Ex = Vs + XS * XK * XP;                                // [2][4 tiles][8 heads][16 rows][EXP] normalised probabilities
  float* Wt = reinterpret_cast<float*>(Ex + 2 * 4 * XH * 16 * EXP);  // [8][8]
  float* nullk = Wt + 64;                                     // [512]
  float* nullv = nullk + XC;                                  // [512]
  float* st_m = nullv + XC;                                   // [XQ][8] row max
  float* st_il = st_m + XQ * XH;                              // [XQ][8] 1 / row sum
  float* st_sn = st_il + XQ * XH;                             // [XQ][8] null-key logit

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = warp >> 2, hp = warp & 3, h0 = hp * 2;       // query tile, head pair
  const int g = lane >> 2, t = lane & 3, id = lane >> 3, l8 = lane & 7;
  const int tiles = (p.nq + XQ - 1) / XQ;
  const int b = blockIdx.x / tiles, q0 = (blockIdx.x - b * tiles) * XQ;
  const bool has_null = p.null_k != nullptr;
  const bf16* qg = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  const bf16* kg = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  const bf16* vg = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs;
  const unsigned char* km = p.key_mask != nullptr ? p.key_mask + (long long)b * p.mask_bs : nullptr;
  const int nchunks = (nk + XK - 1) / XK;
  const float scale = p.qscale;

  // ---- prologue: Q tile + first K chunk in flight, small fp32 tables by plain loads ----
  // cp.async pipeline, XS stages, prefetch distance XS - 1: iteration c waits for chunk c, ONE barrier, then refills the
  // stage chunk c - 1 just vacated with chunk c + XS - 1 (a commit every iteration keeps the group count uniform)
  stage_rows(Qs, qg, p.q_rs, q0, XQ, p.nq);
#pragma unroll
  for (int pc = 0; pc < XS - 1; ++pc) {
    if (pc < nchunks) stage_rows(Ks + pc * XK * XP, kg, p.k_rs, pc * XK, XK, nk);
    cp_async_commit();
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) Wt[i] = p.talk != nullptr ? p.talk[i] : ((i >> 3) == (i & 7) ? 1.f : 0.f);
  for (int i = threadIdx.x; i < XC; i += blockDim.x) {
    nullk[i] = has_null ? p.null_k[i] : 0.f;
    nullv[i] = has_null ? p.null_v[i] : 0.f;
  }
  const int rbase = qt * 16;  // this warp's 16 query rows inside the tile

  // =========================== pass 1: softmax statistics of heads h0, h0 + 1 ===========================
  float m[2][2], l[2][2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) m[hh][0] = m[hh][1] = -FLT_MAX, l[hh][0] = l[hh][1] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<XS - 2>();
    __syncthreads();
    if (c + XS - 1 < nchunks) stage_rows(Ks + ((c + XS - 1) % XS) * XK * XP, kg, p.k_rs, (c + XS - 1) * XK, XK, nk);
    cp_async_commit();
    float s[2][2][4];
    scores2(Qs, Ks + (c % XS) * XK * XP, rbase, h0, l8, id, s);
    bool ok[2][2];  // [8-key block][column]
#pragma unroll
    for (int blk = 0; blk < 2; ++blk)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = c * XK + blk * 8 + 2 * t + e;
        ok[blk][e] = j < nk && (km == nullptr || km[j] != 0);
      }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int r = 0; r < 2; ++r) {  // row g (r = 0) / g + 8 (r = 1)
        float v[4];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int e = 0; e < 2; ++e) v[blk * 2 + e] = ok[blk][e] ? s[hh][blk][r * 2 + e] * scale : -FLT_MAX;
        const float mn = fmaxf(m[hh][r], fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])));
        if (mn > -FLT_MAX) {
          l[hh][r] = l[hh][r] * __expf(m[hh][r] - mn) + __expf(v[0] - mn) + __expf(v[1] - mn) + __expf(v[2] - mn) +
                     __expf(v[3] - mn);
          m[hh][r] = mn;
        }
      }
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with the K stages of pass 1
  // ---- combine the four lanes of a quad, fold in the exact fp32 null-key logit, publish ----
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int h = h0 + hh;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mm = m[hh][r], ll = l[hh][r];
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, mm, o), lo = __shfl_xor_sync(0xffffffffu, ll, o);
        const float mn = fmaxf(mm, mo);
        ll = (mn > -FLT_MAX) ? ll * __expf(mm - mn) + lo * __expf(mo - mn) : 0.f;
        mm = mn;
      }
      float sn = 0.f;
      if (has_null) {
        const bf16* qrow = Qs + (rbase + g + 8 * r) * XP + h * XD + 16 * t;
#pragma unroll
        for (int d = 0; d < 16; ++d) sn = fmaf(__bfloat162float(qrow[d]), nullk[h * XD + 16 * t + d], sn);
        sn += __shfl_xor_sync(0xffffffffu, sn, 1);
        sn += __shfl_xor_sync(0xffffffffu, sn, 2);
        sn *= scale;
        const float mn = fmaxf(mm, sn);
        ll = ll * __expf(mm - mn) + __expf(sn - mn);
        mm = mn;
      }
      m[hh][r] = mm;
      l[hh][r] = ll > 0.f ? 1.0f / ll : 0.f;  // from here on: 1 / row sum
      if (t == 0) {
        const int row = rbase + g + 8 * r;
        st_m[row * XH + h] = mm;
        st_il[row * XH + h] = l[hh][r];
        st_sn[row * XH + h] = sn;
      }
    }
  }
  // first K / V chunks of pass 2 (the stages are free: barrier above)
#pragma unroll
  for (int pc = 0; pc < XS - 1; ++pc) {
    if (pc < nchunks) {
      stage_rows(Ks + pc * XK * XP, kg, p.k_rs, pc * XK, XK, nk);
      stage_rows(Vs + pc * XK * XP, vg, p.v_rs, pc * XK, XK, nk);
    }
    cp_async_commit();
  }

  // =========================== pass 2: normalise, exchange, mix heads, P'V for output heads h0, h0 + 1 ===========================
  float O[2][8][4];
#pragma unroll
  for (int go = 0; go < 2; ++go)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) O[go][nt][0] = O[go][nt][1] = O[go][nt][2] = O[go][nt][3] = 0.f;
  float wmix[2][8];
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<XS - 2>();
    __syncthreads();  // chunk c visible to all; everyone finished chunk c - 1 (its stage is refilled next, and the exchange
                      // buffer of chunk c - 2 may be overwritten)
    if (c + XS - 1 < nchunks) {
      stage_rows(Ks + ((c + XS - 1) % XS) * XK * XP, kg, p.k_rs, (c + XS - 1) * XK, XK, nk);
      stage_rows(Vs + ((c + XS - 1) % XS) * XK * XP, vg, p.v_rs, (c + XS - 1) * XK, XK, nk);
    }
    cp_async_commit();
    if (c == 0) {  // Wt was written before the first barrier of pass 1
#pragma unroll
      for (int go = 0; go < 2; ++go)
#pragma unroll
        for (int h = 0; h < 8; ++h) wmix[go][h] = Wt[(h0 + go) * 8 + h];
    }
    const bf16* Vc = Vs + (c % XS) * XK * XP;
    float s[2][2][4];
    scores2(Qs, Ks + (c % XS) * XK * XP, rbase, h0, l8, id, s);
    // ---- normalised probabilities of my two heads -> exchange buffer (bf16) ----
    bf16* ex = Ex + ((c & 1) * 4 + qt) * XH * 16 * EXP;
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      bool ok[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = c * XK + blk * 8 + 2 * t + e;
        ok[e] = j < nk && (km == nullptr || km[j] != 0);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float p0 = ok[0] ? __expf(s[hh][blk][r * 2] * scale - m[hh][r]) * l[hh][r] : 0.f;
          const float p1 = ok[1] ? __expf(s[hh][blk][r * 2 + 1] * scale - m[hh][r]) * l[hh][r] : 0.f;
          *reinterpret_cast<uint32_t*>(ex + ((h0 + hh) * 16 + g + 8 * r) * EXP + blk * 8 + 2 * t) = pack_bf16x2(p0, p1);
        }
    }
    bar_sync_named(1 + qt, 128);  // the four warps of this query tile have published all 8 heads
    // ---- talking heads: mix the 8 heads for my two output heads, straight into the A fragments of P'V ----
    float mix[2][2][4];
#pragma unroll
    for (int go = 0; go < 2; ++go)
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) mix[go][blk][0] = mix[go][blk][1] = mix[go][blk][2] = mix[go][blk][3] = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
      for (int blk = 0; blk < 2; ++blk)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float2 pv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(ex + (h * 16 + g + 8 * r) * EXP + blk * 8 + 2 * t));
#pragma unroll
          for (int go = 0; go < 2; ++go) {
            mix[go][blk][r * 2] = fmaf(wmix[go][h], pv.x, mix[go][blk][r * 2]);
            mix[go][blk][r * 2 + 1] = fmaf(wmix[go][h], pv.y, mix[go][blk][r * 2 + 1]);
          }
        }
#pragma unroll
    for (int go = 0; go < 2; ++go) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(mix[go][0][0], mix[go][0][1]);  // row g    , keys 2t, 2t+1
      pa[1] = pack_bf16x2(mix[go][0][2], mix[go][0][3]);  // row g + 8
      pa[2] = pack_bf16x2(mix[go][1][0], mix[go][1][1]);  // row g    , keys 8 + 2t ..
      pa[3] = pack_bf16x2(mix[go][1][2], mix[go][1][3]);  // row g + 8
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {
        uint32_t vb[4];
        ldsm4t(vb, Vc + (l8 + 8 * (id & 1)) * XP + (h0 + go) * XD + ntp * 16 + 8 * (id >> 1));
        mma16816(O[go][ntp * 2], pa, vb[0], vb[1]);
        mma16816(O[go][ntp * 2 + 1], pa, vb[2], vb[3]);
      }
    }
  }

  // ---- null value (fp32), store ----
  bf16* ob = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs;
#pragma unroll
  for (int go = 0; go < 2; ++go) {
    const int hg = h0 + go;
    float pm[2] = {0.f, 0.f};
    if (has_null) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = rbase + g + 8 * r;
#pragma unroll
        for (int h = 0; h < 8; ++h)
          pm[r] = fmaf(wmix[go][h], __expf(st_sn[row * XH + h] - st_m[row * XH + h]) * st_il[row * XH + h], pm[r]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = hg * XD + nt * 8 + 2 * t;
      const float nv0 = nullv[d], nv1 = nullv[d + 1];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int qrow = q0 + rbase + g + 8 * r;
        if (qrow < p.nq)
          *reinterpret_cast<uint32_t*>(ob + (long long)qrow * p.o_rs + d) =
              pack_bf16x2(fmaf(pm[r], nv0, O[go][nt][r * 2]), fmaf(pm[r], nv1, O[go][nt][r * 2 + 1]));
      }
    }
  }
}

}  // namespace

// Returns NUWA_ERR_INVALID outside the envelope (H = 8, dh = 64, no additive bias / per-head scale): callers fall back to
// attention_mma.cu / the generic kernel.
int attn_dense_x64(const AttnParams& p, int nk, cudaStream_t stream) {
  if (p.H != XH || p.dh != XD || p.bias != nullptr || p.head_scale != nullptr || nk <= 0 || p.nq < 16) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.k_rs % 8) || (p.v_rs % 8) || (p.q_bs % 8) || (p.k_bs % 8) || (p.v_bs % 8) || (p.o_rs & 1))
    return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.k) & 15) || (reinterpret_cast<uintptr_t>(p.v) & 15))
    return NUWA_ERR_INVALID;
  const size_t smem = (size_t)(XQ + 2 * XS * XK) * XP * 2 + (size_t)2 * 4 * XH * 16 * EXP * 2 + (64 + 2 * XC + 3 * XQ * XH) * sizeof(float);
  static const cudaError_t attr_rc =   // one-time, thread-safe static initialisation (smem is a compile-time constant)
      cudaFuncSetAttribute(attn_dense_x64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr_rc != cudaSuccess) return NUWA_ERR_CUDA;
  const int grid = p.B * ((p.nq + XQ - 1) / XQ);
  attn_dense_x64_kernel<<<grid, XT, smem, stream>>>(p, nk);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
