// Fused attention cores: neighbourhood gather / dense keys + QK^T + mask + fp32 softmax +
// talking-heads mix across heads + PV, one kernel, nothing materialised in HBM.
//
// One warp owns QPW query tokens with ALL heads (the post-softmax talking-heads 1x1 conv mixes
// probabilities across heads, nuwa_pytorch.py:556-558 / :372 / :889, so every head's row must exist
// before any PV).  The 32 lanes split the H*dh channels (CPL per lane, 32/H lanes per head); scores
// live in shared memory as [QPW][H][J]; K/V rows are read with 16-byte vector loads straight from
// the bf16 q|k|v projection buffer (neighbour reuse is served by L1/L2).
//
// Key generators:
//   MODE_3DNA   Sparse3DNA (nuwa_pytorch.py:459-613): bos key + (kt,kh,kw) dilated window, causal or
//               centred; out-of-grid slots are masked, in-grid slots beyond the supplied tokens are
//               visible ZERO keys; query 0 (bos) copies its own value (:608).
//   MODE_DENSE  Attention (nuwa_pytorch.py:315-379): optional learned null key/value (fp32), key mask,
//               optional additive bias + per-head logit scale (VQGanAttention, vqgan_vae.py:275-279).
//   MODE_X2DNA  SparseCross2DNA non-bos queries (nuwa_pytorch.py:851-895): null + k x k window at the
//               query's own (y,x) in every context frame.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

enum { MODE_3DNA = 0, MODE_DENSE = 1, MODE_X2DNA = 2 };
enum { KEY_NORMAL = 0, KEY_MASKED = 1, KEY_ZERO = 2, KEY_NULL = 3 };

template <int CPL>
__device__ __forceinline__ void load_row(const bf16* p, float (&out)[CPL]) {
  if constexpr (CPL >= 8) {
#pragma unroll
    for (int i = 0; i < CPL / 8; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      out[i * 8 + 0] = a.x; out[i * 8 + 1] = a.y; out[i * 8 + 2] = b.x; out[i * 8 + 3] = b.y;
      out[i * 8 + 4] = c.x; out[i * 8 + 5] = c.y; out[i * 8 + 6] = d.x; out[i * 8 + 7] = d.y;
    }
  } else if constexpr (CPL == 4) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
  } else if constexpr (CPL == 2) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    float2 a = unpack_bf16x2(u);
    out[0] = a.x; out[1] = a.y;
  } else {
    out[0] = __bfloat162float(p[0]);
  }
}
template <int CPL>
__device__ __forceinline__ void load_row_f32(const float* p, float (&out)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; ++i) out[i] = __ldg(p + i);
}
template <int CPL>
__device__ __forceinline__ void store_row(bf16* p, const float (&v)[CPL]) {
  if constexpr (CPL >= 8) {
#pragma unroll
    for (int i = 0; i < CPL / 8; ++i) {
      uint4 u;
      u.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
      u.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
      u.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
      u.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
      reinterpret_cast<uint4*>(p)[i] = u;
    }
  } else if constexpr (CPL == 4) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  } else if constexpr (CPL == 2) {
    *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]);
  } else {
    p[0] = __float2bfloat16(v[0]);
  }
}

// key slot j of query position t (absolute sequence index) -> kind and row in the K/V buffers
template <int MODE>
__device__ __forceinline__ int key_of(const AttnParams& p, int b, int t, int nv, int j, int& row) {
  if constexpr (MODE == MODE_3DNA) {
    if (j == 0) {
      row = 0;  // bos key / value
      return KEY_NORMAL;
    }
    const int jj = j - 1;
    const int c = jj % p.kw, bq = (jj / p.kw) % p.kh, a = jj / (p.kw * p.kh);
    const int T = p.fmap * p.fmap;
    const int vt = t - 1;
    const int f = vt / T, y = (vt % T) / p.fmap, x = vt % p.fmap;
    const int pf = p.dt * (p.kt - 1) / 2, ph = p.dh_ * (p.kh - 1) / 2, pw = p.dw * (p.kw - 1) / 2;
    const int Pf = p.causal ? 2 * pf : pf, Ph = p.causal ? 2 * ph : ph, Pw = p.causal ? 2 * pw : pw;
    const int ff = f + a * p.dt - Pf, yy = y + bq * p.dh_ - Ph, xx = x + c * p.dw - Pw;
    if (ff < 0 || ff >= p.max_frames || yy < 0 || yy >= p.fmap || xx < 0 || xx >= p.fmap) return KEY_MASKED;
    const int idx = (ff * p.fmap + yy) * p.fmap + xx;
    if (idx >= nv) return KEY_ZERO;  // zero-padded position: visible, contributes exp(0 - max), no value
    row = 1 + idx;
    return KEY_NORMAL;
  } else if constexpr (MODE == MODE_DENSE) {
    int jj = j;
    if (p.null_k != nullptr) {
      if (j == 0) return KEY_NULL;
      jj = j - 1;
    }
    if (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + jj] == 0) return KEY_MASKED;
    row = jj;
    return KEY_NORMAL;
  } else {
    if (j == 0) return KEY_NULL;
    const int jj = j - 1;
    const int J2 = p.ck * p.ck;
    const int f = jj / J2, w = jj % J2;
    const int a = w / p.ck, c = w % p.ck;
    const int T = p.fmap * p.fmap;
    const int i = (t - 1) % T;
    const int y = i / p.fmap, x = i % p.fmap;
    const int pad = p.cdil * (p.ck - 1) / 2;
    const int yy = y + a * p.cdil - pad, xx = x + c * p.cdil - pad;
    if (yy < 0 || yy >= p.fmap || xx < 0 || xx >= p.fmap) return KEY_MASKED;
    const int idx = f * T + yy * p.fmap + xx;
    if (p.key_mask != nullptr && p.key_mask[(long long)b * p.mask_bs + idx] == 0) return KEY_MASKED;
    row = idx;
    return KEY_NORMAL;
  }
}

template <int MODE, int CPL, int QPW>
__global__ void __launch_bounds__(128) attn_kernel(const AttnParams p) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int lph = 32 / p.H;            // lanes per head
  const int h = lane / lph;            // head of this lane
  const int sub = lane - h * lph;      // lane index inside the head group
  const int J = p.jmax;                // key slots per query
  float* S = smem_f + (size_t)warp * QPW * p.H * J;   // [QPW][H][J]
  float* Wt = smem_f + (size_t)warps_per_cta * QPW * p.H * J;  // [H][H] talking-heads matrix
  int* keys = reinterpret_cast<int*>(Wt + p.H * p.H) + warp * J;  // per-warp key list: (kind << 28) | row
  if (p.talk != nullptr) {
    for (int i = threadIdx.x; i < p.H * p.H; i += blockDim.x) Wt[i] = p.talk[i];
  }
  __syncthreads();

  const int groups_per_b = (p.nq + QPW - 1) / QPW;
  const int gidx = blockIdx.x * warps_per_cta + warp;
  if (gidx >= p.B * groups_per_b) return;
  const int b = gidx / groups_per_b;
  const int q0 = (gidx - b * groups_per_b) * QPW;  // first local query index of this warp
  const int ch = lane * CPL;                        // first channel of this lane

  const bf16* kb = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  const bf16* vb = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs;
  const bf16* qb = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  bf16* ob = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs;

  float qf[QPW][CPL];
  bool qok[QPW];
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    qok[qi] = (q0 + qi) < p.nq;
    if (qok[qi]) {
      load_row<CPL>(qb + (long long)(q0 + qi) * p.q_rs + ch, qf[qi]);
#pragma unroll
      for (int c = 0; c < CPL; ++c) qf[qi][c] *= p.qscale;
    } else {
#pragma unroll
      for (int c = 0; c < CPL; ++c) qf[qi][c] = 0.f;
    }
  }
  const float hscale = p.head_scale != nullptr ? p.head_scale[h] : 1.0f;

  // Sparse3DNA: the bos query attends only to itself -> output = its own value row
  const int t0 = p.t0_ptr != nullptr ? __ldg(p.t0_ptr) : p.t0;  // device-side position for graph-replayed decode steps
  const int nv = (p.t0_ptr != nullptr && MODE == MODE_3DNA) ? t0 : p.nv;
  if (MODE == MODE_3DNA && (t0 + q0) == 0) {
    float vv[CPL];
    load_row<CPL>(vb + ch, vv);
    store_row<CPL>(ob + (long long)q0 * p.o_rs + ch, vv);
    qok[0] = false;  // (QPW == 1 for this mode)
    if (QPW == 1) return;
  }

  // key list once per warp (the index arithmetic is ~40 integer ops per slot: do it on one lane per slot instead of
  // on all 32 lanes per slot, twice).  Key kind is query dependent only for the gather modes (QPW == 1 there).
  for (int j = lane; j < J; j += 32) {
    int row = 0;
    const int kind = key_of<MODE>(p, b, t0 + q0, nv, j, row);
    keys[j] = (kind << 28) | row;
  }
  __syncwarp();

  // ---------------- scores ----------------
  for (int j = 0; j < J; ++j) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF;
    const int kind = kj >> 28;
    float part[QPW];
    if (kind == KEY_NORMAL || kind == KEY_NULL) {
      float kf[CPL];
      if (kind == KEY_NULL) load_row_f32<CPL>(p.null_k + ch, kf);
      else load_row<CPL>(kb + (long long)row * p.k_rs + ch, kf);
#pragma unroll
      for (int qi = 0; qi < QPW; ++qi) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) s = fmaf(qf[qi][c], kf[c], s);
        part[qi] = s;
      }
      for (int o = lph >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi) part[qi] += __shfl_xor_sync(0xffffffffu, part[qi], o);
      }
    } else {
#pragma unroll
      for (int qi = 0; qi < QPW; ++qi) part[qi] = (kind == KEY_MASKED) ? -FLT_MAX : 0.f;
    }
    if (sub == 0) {
#pragma unroll
      for (int qi = 0; qi < QPW; ++qi) {
        float s = part[qi];
        if (kind != KEY_MASKED) {
          s *= hscale;
          if (p.bias != nullptr)
            s += p.bias[((long long)h * p.bias_nq + (t0 + q0 + qi < p.bias_nq ? t0 + q0 + qi : 0)) * p.bias_nk + j];
        }
        S[(qi * p.H + h) * J + j] = s;
      }
    }
  }
  __syncwarp();

  // ---------------- fp32 softmax per (query, head) row ----------------
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    float* row_s = S + (qi * p.H + h) * J;
    float m = -FLT_MAX;
    for (int j = sub; j < J; j += lph) m = fmaxf(m, row_s[j]);
    for (int o = lph >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = sub; j < J; j += lph) {
      const float e = __expf(row_s[j] - m);
      row_s[j] = e;
      sum += e;
    }
    for (int o = lph >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = sub; j < J; j += lph) row_s[j] *= inv;
  }
  __syncwarp();

  // ---------------- talking heads: P'[g][j] = sum_h W[g][h] P[h][j] ----------------
  if (p.talk != nullptr) {
    for (int qi = 0; qi < QPW; ++qi) {
      float* base = S + (size_t)qi * p.H * J;
      for (int j = lane; j < J; j += 32) {
        float pin[32];
        for (int hh = 0; hh < p.H; ++hh) pin[hh] = base[hh * J + j];
        for (int g = 0; g < p.H; ++g) {
          float acc = 0.f;
          for (int hh = 0; hh < p.H; ++hh) acc = fmaf(Wt[g * p.H + hh], pin[hh], acc);
          base[g * J + j] = acc;
        }
      }
    }
    __syncwarp();
  }

  // ---------------- PV ----------------
  float acc[QPW][CPL];
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi)
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[qi][c] = 0.f;
  for (int j = 0; j < J; ++j) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF;
    const int kind = kj >> 28;
    if (kind == KEY_MASKED || kind == KEY_ZERO) continue;
    float vf[CPL];
    if (kind == KEY_NULL) load_row_f32<CPL>(p.null_v + ch, vf);
    else load_row<CPL>(vb + (long long)row * p.v_rs + ch, vf);
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const float pj = S[(qi * p.H + h) * J + j];
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[qi][c] = fmaf(pj, vf[c], acc[qi][c]);
    }
  }
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi)
    if (qok[qi]) store_row<CPL>(ob + (long long)(q0 + qi) * p.o_rs + ch, acc[qi]);
}

template <int MODE, int QPW>
static int launch_attn(const AttnParams& p, cudaStream_t stream) {
  const int inner = p.H * p.dh;
  if (inner % 32 != 0 || (32 % p.H) != 0 || p.H > 32) return NUWA_ERR_INVALID;
  const int cpl = inner / 32;
  const int warps = 4;
  const size_t smem = ((size_t)warps * QPW * p.H * p.jmax + p.H * p.H + (size_t)warps * p.jmax) * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  const int groups = p.B * ((p.nq + QPW - 1) / QPW);
  if (groups <= 0) return NUWA_OK;
  const int grid = (groups + warps - 1) / warps;
#define NUWA_LAUNCH_ATTN(CPL)                                                                                  \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      cudaFuncSetAttribute(attn_kernel<MODE, CPL, QPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    attn_kernel<MODE, CPL, QPW><<<grid, warps * 32, smem, stream>>>(p);                                        \
  } while (0)
  switch (cpl) {
    case 16: NUWA_LAUNCH_ATTN(16); break;
    case 8: NUWA_LAUNCH_ATTN(8); break;
    case 4: NUWA_LAUNCH_ATTN(4); break;
    case 2: NUWA_LAUNCH_ATTN(2); break;
    case 1: NUWA_LAUNCH_ATTN(1); break;
    default: return NUWA_ERR_INVALID;
  }
#undef NUWA_LAUNCH_ATTN
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

int attn_decode_sparse3dna(const AttnParams& p, cudaStream_t s);
int attn_decode_dense(const AttnParams& p, cudaStream_t s);
int attn_decode_cross2dna(const AttnParams& p, cudaStream_t s);

int attn_sparse3dna(const AttnParams& p, cudaStream_t s) {
  if (p.nq == 1) {
    const int rc = attn_decode_sparse3dna(p, s);
    if (rc != NUWA_ERR_INVALID) return rc;
  }
  return launch_attn<MODE_3DNA, 1>(p, s);
}
int attn_dense(const AttnParams& p, cudaStream_t s) {
  // several queries per warp amortise every K/V row load; single-query decode steps use QPW=1
  if (p.nq >= 4) return launch_attn<MODE_DENSE, 4>(p, s);
  if (p.nq == 1) {
    const int rc = attn_decode_dense(p, s);
    if (rc != NUWA_ERR_INVALID) return rc;
  }
  return launch_attn<MODE_DENSE, 1>(p, s);
}
int attn_cross2dna(const AttnParams& p, cudaStream_t s) {
  if (p.nq == 1) {
    const int rc = attn_decode_cross2dna(p, s);
    if (rc != NUWA_ERR_INVALID) return rc;
  }
  return launch_attn<MODE_X2DNA, 1>(p, s);
}

}  // namespace nuwa

// ================================================================================================
// Decode-step attention (one query position per sample): the single-query problem has no query
// parallelism, so the KEYS are spread over the lanes instead -- lane j scores key j with the whole
// head vector in registers, the softmax is a warp reduction, the talking-heads mix goes through
// shared memory, and PV gives every lane dh/32 output channels with coalesced V row reads.
// grid = B, block = 32*H (warp = head).  Used by generate() (Sparse3DNA / dense / cross-2DNA).
// ================================================================================================
namespace nuwa {

template <int MODE, int DH>
__global__ void __launch_bounds__(256) attn_decode_kernel(const AttnParams p) {
  extern __shared__ float smem_d[];
  const int H = p.H, J = p.jmax;
  float* P = smem_d;                                  // [H][J]
  float* Wt = P + (size_t)H * J;                      // [H][H]
  int* keys = reinterpret_cast<int*>(Wt + H * H);     // [J]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int t0 = p.t0_ptr != nullptr ? __ldg(p.t0_ptr) : p.t0;
  const int nv = (p.t0_ptr != nullptr && MODE == MODE_3DNA) ? t0 : p.nv;
  const bf16* kb = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  const bf16* vb = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs;
  const bf16* qb = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  bf16* ob = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs;
  constexpr int DPL = DH / 32;  // output channels per lane
  if (MODE == MODE_3DNA && t0 == 0) {  // bos attends only to itself
    for (int c = threadIdx.x; c < H * DH; c += blockDim.x) ob[c] = vb[c];
    return;
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    int row = 0;
    const int kind = key_of<MODE>(p, b, t0, nv, j, row);
    keys[j] = (kind << 28) | row;
  }
  if (p.talk != nullptr)
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) Wt[i] = p.talk[i];
  __syncthreads();
  // ---- scores: lane j <-> key j ----
  float qf[DH];
  {
    const bf16* qh = qb + w * DH;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(qh) + i);
      const float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      qf[i * 8 + 0] = a.x; qf[i * 8 + 1] = a.y; qf[i * 8 + 2] = bq.x; qf[i * 8 + 3] = bq.y;
      qf[i * 8 + 4] = c.x; qf[i * 8 + 5] = c.y; qf[i * 8 + 6] = d.x; qf[i * 8 + 7] = d.y;
    }
  }
  const float hs = p.qscale * (p.head_scale != nullptr ? p.head_scale[w] : 1.0f);
  float* Pw = P + (size_t)w * J;
  float m = -FLT_MAX;
  for (int j = lane; j < J; j += 32) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    float s;
    if (kind == KEY_NORMAL) {
      const bf16* kr = kb + (long long)row * p.k_rs + w * DH;
      s = 0.f;
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(kr) + i);
        const float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        s += qf[i * 8 + 0] * a.x + qf[i * 8 + 1] * a.y + qf[i * 8 + 2] * bq.x + qf[i * 8 + 3] * bq.y +
             qf[i * 8 + 4] * c.x + qf[i * 8 + 5] * c.y + qf[i * 8 + 6] * d.x + qf[i * 8 + 7] * d.y;
      }
      s *= hs;
    } else if (kind == KEY_NULL) {
      const float* nk = p.null_k + w * DH;
      s = 0.f;
#pragma unroll
      for (int i = 0; i < DH; ++i) s += qf[i] * __ldg(nk + i);
      s *= hs;
    } else {
      s = (kind == KEY_MASKED) ? -FLT_MAX : 0.f;
    }
    if (kind != KEY_MASKED && p.bias != nullptr) s += p.bias[((long long)w * p.bias_nq + (t0 < p.bias_nq ? t0 : 0)) * p.bias_nk + j];
    Pw[j] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float l = 0.f;
  for (int j = lane; j < J; j += 32) {
    const float e = __expf(Pw[j] - m);
    Pw[j] = e;
    l += e;
  }
  l = warp_sum(l);
  const float inv = 1.0f / l;
  for (int j = lane; j < J; j += 32) Pw[j] *= inv;
  __syncthreads();
  // ---- talking heads ----
  if (p.talk != nullptr) {
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      float pin[8];
      for (int h = 0; h < H; ++h) pin[h] = P[(size_t)h * J + j];
      for (int g = 0; g < H; ++g) {
        float a = 0.f;
        for (int h = 0; h < H; ++h) a = fmaf(Wt[g * H + h], pin[h], a);
        P[(size_t)g * J + j] = a;
      }
    }
    __syncthreads();
  }
  // ---- PV: the keys stay spread over the lanes (every lane streams whole V rows of ITS keys with independent
  //      16-byte loads -> full memory-level parallelism), then a shared-memory reduction folds the 32 partial
  //      head vectors and gives lane l the output channels [l*DPL, (l+1)*DPL) ----
  float acc[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  for (int j = lane; j < J; j += 32) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    if (kind == KEY_MASKED || kind == KEY_ZERO) continue;
    const float pj = Pw[j];
    if (kind == KEY_NULL) {
      const float* nvp = p.null_v + w * DH;
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = fmaf(pj, __ldg(nvp + c), acc[c]);
    } else {
      const bf16* vr = vb + (long long)row * p.v_rs + w * DH;
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(vr) + i);
        const float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[i * 8 + 0] = fmaf(pj, a.x, acc[i * 8 + 0]); acc[i * 8 + 1] = fmaf(pj, a.y, acc[i * 8 + 1]);
        acc[i * 8 + 2] = fmaf(pj, bq.x, acc[i * 8 + 2]); acc[i * 8 + 3] = fmaf(pj, bq.y, acc[i * 8 + 3]);
        acc[i * 8 + 4] = fmaf(pj, c2.x, acc[i * 8 + 4]); acc[i * 8 + 5] = fmaf(pj, c2.y, acc[i * 8 + 5]);
        acc[i * 8 + 6] = fmaf(pj, d.x, acc[i * 8 + 6]); acc[i * 8 + 7] = fmaf(pj, d.y, acc[i * 8 + 7]);
      }
    }
  }
  float* red = reinterpret_cast<float*>(keys + J) + (size_t)w * 32 * (DH + 1);  // [32 lanes][DH + 1]
#pragma unroll
  for (int c = 0; c < DH; ++c) red[lane * (DH + 1) + c] = acc[c];
  __syncwarp();
  float outv[DPL];
#pragma unroll
  for (int c = 0; c < DPL; ++c) {
    float sum = 0.f;
    for (int l = 0; l < 32; ++l) sum += red[l * (DH + 1) + lane * DPL + c];
    outv[c] = sum;
  }
  store_row<DPL>(ob + w * DH + lane * DPL, outv);
}

template <int MODE>
static int launch_attn_decode(const AttnParams& p, cudaStream_t stream) {
  if (p.nq != 1 || p.H > 8 || (p.dh != 64 && p.dh != 32)) return NUWA_ERR_INVALID;
  if ((p.k_rs % 8) || (p.q_rs % 8) || (p.k_bs % 8) || (p.q_bs % 8)) return NUWA_ERR_INVALID;
  const size_t smem = ((size_t)p.H * p.jmax + p.H * p.H + p.jmax + (size_t)p.H * 32 * (p.dh + 1)) * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  if (p.dh == 64) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_decode_kernel<MODE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_decode_kernel<MODE, 64><<<p.B, 32 * p.H, smem, stream>>>(p);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_decode_kernel<MODE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_decode_kernel<MODE, 32><<<p.B, 32 * p.H, smem, stream>>>(p);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

int attn_decode_sparse3dna(const AttnParams& p, cudaStream_t s) { return launch_attn_decode<MODE_3DNA>(p, s); }
int attn_decode_dense(const AttnParams& p, cudaStream_t s) { return launch_attn_decode<MODE_DENSE>(p, s); }
int attn_decode_cross2dna(const AttnParams& p, cudaStream_t s) { return launch_attn_decode<MODE_X2DNA>(p, s); }

}  // namespace nuwa

// ================================================================================================
// Backward of the gather attentions (Sparse3DNA / SparseCross2DNA), see backward.cu for the shared row kernel.
//   gather_scores : per query, S[h][j] = qscale q.k_j (masked slots -FLT_MAX, zero keys 0) and dP'[h][j] = dO.v_j
//   gather_dq     : dq = sum_j dS[h][j] k_j               (dS already carries qscale)
//   gather_dkdv   : key-centric -- every key row sums over the queries whose window contains it (inverse of key_of),
//                   no atomics, deterministic
//   first_key     : slot 0 (bos key of Sparse3DNA / null key of SparseCross2DNA) is seen by every query: column reduction
// Score / probability tensors are [B][H][nq][jp] with nq = the non-bos queries (absolute positions t0 .. t0+nq-1).
// ================================================================================================
namespace nuwa {

template <int MODE, int CPL>
__global__ void __launch_bounds__(128) gather_scores_kernel(const AttnParams p, const bf16* __restrict__ dO, long long do_bs,
                                                            int do_rs, float* __restrict__ S, float* __restrict__ dPp, int jp) {
  extern __shared__ float smem_g[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int lph = 32 / p.H, h = lane / lph, sub = lane - h * lph;
  const int J = p.jmax;
  float* Ss = smem_g + (size_t)warp * (2 * p.H * J + J);  // [H][J] scores, [H][J] dP', [J] keys
  float* Ds = Ss + p.H * J;
  int* keys = reinterpret_cast<int*>(Ds + p.H * J);
  const long long gidx = blockIdx.x * (long long)wpb + warp;
  if (gidx >= (long long)p.B * p.nq) return;
  const int b = (int)(gidx / p.nq), ql = (int)(gidx - (long long)b * p.nq);
  const int t = p.t0 + ql;
  const int ch = lane * CPL;
  const bf16* kb = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  const bf16* vb = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs;
  float qf[CPL], df[CPL];
  load_row<CPL>(reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs + (long long)ql * p.q_rs + ch, qf);
  load_row<CPL>(dO + (long long)b * do_bs + (long long)ql * do_rs + ch, df);
#pragma unroll
  for (int c = 0; c < CPL; ++c) qf[c] *= p.qscale;
  for (int j = lane; j < J; j += 32) {
    int row = 0;
    const int kind = key_of<MODE>(p, b, t, p.nv, j, row);
    keys[j] = (kind << 28) | row;
  }
  __syncwarp();
  for (int j = 0; j < J; ++j) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    float s, d = 0.f;
    if (kind == KEY_NORMAL || kind == KEY_NULL) {
      float kf[CPL], vf[CPL];
      if (kind == KEY_NULL) {
        load_row_f32<CPL>(p.null_k + ch, kf);
        load_row_f32<CPL>(p.null_v + ch, vf);
      } else {
        load_row<CPL>(kb + (long long)row * p.k_rs + ch, kf);
        load_row<CPL>(vb + (long long)row * p.v_rs + ch, vf);
      }
      s = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        s = fmaf(qf[c], kf[c], s);
        d = fmaf(df[c], vf[c], d);
      }
      for (int o = lph >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
      }
    } else {
      s = (kind == KEY_MASKED) ? -FLT_MAX : 0.f;
    }
    if (sub == 0) {
      Ss[h * J + j] = s;
      Ds[h * J + j] = d;
    }
  }
  __syncwarp();
  for (int hh = 0; hh < p.H; ++hh) {
    const long long o = (((long long)b * p.H + hh) * p.nq + ql) * jp;
    for (int j = lane; j < J; j += 32) {
      S[o + j] = Ss[hh * J + j];
      dPp[o + j] = Ds[hh * J + j];
    }
  }
}

template <int MODE, int CPL>
__global__ void __launch_bounds__(128) gather_dq_kernel(const AttnParams p, const bf16* __restrict__ dS, int jp,
                                                        bf16* __restrict__ dq, long long dq_bs, int dq_rs) {
  extern __shared__ float smem_g[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int lph = 32 / p.H, h = lane / lph;
  const int J = p.jmax;
  float* Ds = smem_g + (size_t)warp * (p.H * J + J);
  int* keys = reinterpret_cast<int*>(Ds + p.H * J);
  const long long gidx = blockIdx.x * (long long)wpb + warp;
  if (gidx >= (long long)p.B * p.nq) return;
  const int b = (int)(gidx / p.nq), ql = (int)(gidx - (long long)b * p.nq);
  const int t = p.t0 + ql;
  const int ch = lane * CPL;
  const bf16* kb = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  for (int j = lane; j < J; j += 32) {
    int row = 0;
    const int kind = key_of<MODE>(p, b, t, p.nv, j, row);
    keys[j] = (kind << 28) | row;
  }
  for (int hh = 0; hh < p.H; ++hh) {
    const long long o = (((long long)b * p.H + hh) * p.nq + ql) * jp;
    for (int j = lane; j < J; j += 32) Ds[hh * J + j] = __bfloat162float(dS[o + j]);
  }
  __syncwarp();
  float acc[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
  for (int j = 0; j < J; ++j) {
    const int kj = keys[j];
    const int row = kj & 0x0FFFFFFF, kind = kj >> 28;
    if (kind != KEY_NORMAL && kind != KEY_NULL) continue;
    float kf[CPL];
    if (kind == KEY_NULL) load_row_f32<CPL>(p.null_k + ch, kf);
    else load_row<CPL>(kb + (long long)row * p.k_rs + ch, kf);
    const float w = Ds[h * J + j];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = fmaf(w, kf[c], acc[c]);
  }
  store_row<CPL>(dq + (long long)b * dq_bs + (long long)ql * dq_rs + ch, acc);
}

// Sparse3DNA key-centric pass: key row r (video token r-1) <- queries (f,y,x) = key - offset*dil + P  (inverse of key_of)
template <int CPL>
__global__ void __launch_bounds__(128)
gather_dkdv_3dna_kernel(const AttnParams p, const bf16* __restrict__ dO, long long do_bs, int do_rs,
                        const bf16* __restrict__ dS, const bf16* __restrict__ Pp, int jp, bf16* __restrict__ dk,
                        bf16* __restrict__ dv, long long dkv_bs, int dkv_rs) {
  extern __shared__ float smem_g[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int lph = 32 / p.H, h = lane / lph;
  const int J = p.jmax;
  int* inv = reinterpret_cast<int*>(smem_g) + (size_t)warp * J;  // (query local index << 8) | slot, or -1
  const long long gidx = blockIdx.x * (long long)wpb + warp;
  if (gidx >= (long long)p.B * p.nv) return;
  const int b = (int)(gidx / p.nv), kidx = (int)(gidx - (long long)b * p.nv);  // video token index of this key
  const int T = p.fmap * p.fmap;
  const int kf_ = kidx / T, ky = (kidx % T) / p.fmap, kx = kidx % p.fmap;
  const int pf = p.dt * (p.kt - 1) / 2, ph = p.dh_ * (p.kh - 1) / 2, pw = p.dw * (p.kw - 1) / 2;
  const int Pf = p.causal ? 2 * pf : pf, Ph = p.causal ? 2 * ph : ph, Pw = p.causal ? 2 * pw : pw;
  for (int jj = lane; jj < J - 1; jj += 32) {
    const int c = jj % p.kw, bq = (jj / p.kw) % p.kh, a = jj / (p.kw * p.kh);
    const int qf_ = kf_ - a * p.dt + Pf, qy = ky - bq * p.dh_ + Ph, qx = kx - c * p.dw + Pw;
    int e = -1;
    if (qf_ >= 0 && qf_ < p.max_frames && qy >= 0 && qy < p.fmap && qx >= 0 && qx < p.fmap) {
      const int qidx = (qf_ * p.fmap + qy) * p.fmap + qx;  // video token index of the query; its position is 1 + qidx
      const int ql = 1 + qidx - p.t0;
      if (ql >= 0 && ql < p.nq) e = (ql << 8) | (jj + 1);
    }
    inv[jj] = e;
  }
  __syncwarp();
  const int ch = lane * CPL;
  const bf16* qb = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  const bf16* dob = dO + (long long)b * do_bs;
  float ak[CPL], av[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) ak[c] = av[c] = 0.f;
  const long long hb = ((long long)b * p.H + h) * p.nq;
  for (int jj = 0; jj < J - 1; ++jj) {
    const int e = inv[jj];
    if (e < 0) continue;
    const int ql = e >> 8, j = e & 255;
    const float ws = __bfloat162float(dS[(hb + ql) * jp + j]);
    const float wp = __bfloat162float(Pp[(hb + ql) * jp + j]);
    float qv[CPL], dv_[CPL];
    load_row<CPL>(qb + (long long)ql * p.q_rs + ch, qv);
    load_row<CPL>(dob + (long long)ql * do_rs + ch, dv_);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      ak[c] = fmaf(ws, qv[c], ak[c]);
      av[c] = fmaf(wp, dv_[c], av[c]);
    }
  }
  const long long o = (long long)b * dkv_bs + (long long)(1 + kidx) * dkv_rs + ch;
  store_row<CPL>(dk + o, ak);
  store_row<CPL>(dv + o, av);
}

// slot-0 key (seen by every query): out_k[b][c] += sum_q dS[b][h][q][0] Q[b][q][c] ; out_v[b][c] += sum_q P'[..][0] dO[b][q][c]
// grid (chunks, B), block = inner threads; ok_bs = 0 accumulates all samples into one vector (learned null key).
__global__ void __launch_bounds__(1024)
first_key_kernel(const bf16* __restrict__ q, long long q_bs, int q_rs, const bf16* __restrict__ dO, long long do_bs, int do_rs,
                 const bf16* __restrict__ dS, const bf16* __restrict__ Pp, int jp, int H, int dh, int nq, int chunk,
                 float* __restrict__ out_k, float* __restrict__ out_v, long long ok_bs) {
  // The slot-0 weights of the chunk's queries are staged in shared memory first (they are one strided 2-byte load per
  // (head, query): read inside the accumulation loop they put two dependent L2 round trips on every iteration), then
  // the loop streams the q / dO rows (coalesced across the channel threads) with 8 independent loads in flight.
  extern __shared__ float fk_w[];  // [2][H][chunk]
  const int b = blockIdx.y, c = threadIdx.x, h = c / dh;
  const int q0 = blockIdx.x * chunk, q1 = min(nq, q0 + chunk);
  float* ws_s = fk_w;
  float* wp_s = fk_w + H * chunk;
  for (int i = threadIdx.x; i < H * chunk; i += blockDim.x) {
    const int hh = i / chunk, ql = q0 + (i - hh * chunk);
    float a = 0.f, p2 = 0.f;
    if (ql < nq) {
      const long long o = (((long long)b * H + hh) * nq + ql) * jp;
      a = __bfloat162float(dS[o]);
      p2 = __bfloat162float(Pp[o]);
    }
    ws_s[i] = a;
    wp_s[i] = p2;
  }
  __syncthreads();
  float ak = 0.f, av = 0.f;
  const bf16* qp = q + (long long)b * q_bs + (long long)q0 * q_rs + c;
  const bf16* dp = dO + (long long)b * do_bs + (long long)q0 * do_rs + c;
  const int n = q1 - q0;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    float qv[8], dv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      qv[u] = __bfloat162float(qp[(long long)(i + u) * q_rs]);
      dv[u] = __bfloat162float(dp[(long long)(i + u) * do_rs]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      ak = fmaf(ws_s[h * chunk + i + u], qv[u], ak);
      av = fmaf(wp_s[h * chunk + i + u], dv[u], av);
    }
  }
  for (; i < n; ++i) {
    ak = fmaf(ws_s[h * chunk + i], __bfloat162float(qp[(long long)i * q_rs]), ak);
    av = fmaf(wp_s[h * chunk + i], __bfloat162float(dp[(long long)i * do_rs]), av);
  }
  atomicAdd(out_k + (long long)b * ok_bs + c, ak);
  atomicAdd(out_v + (long long)b * ok_bs + c, av);
}
// Sparse3DNA row 0 of d(q|k|v): dq = 0 (the bos query copies its value, :608), dk = tmp_k, dv = tmp_v + dO[bos]
__global__ void __launch_bounds__(256)
first_key_finalize_kernel(const float* __restrict__ tmp_k, const float* __restrict__ tmp_v, const bf16* __restrict__ dO_bos,
                          long long do_bs, bf16* __restrict__ dqkv, long long dqkv_bs, int inner, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * inner) return;
  const int b = i / inner, c = i - b * inner;
  bf16* o = dqkv + (long long)b * dqkv_bs;
  o[c] = __float2bfloat16(0.f);
  o[inner + c] = __float2bfloat16(tmp_k[(long long)b * inner + c]);
  o[2 * inner + c] = __float2bfloat16(tmp_v[(long long)b * inner + c] + __bfloat162float(dO_bos[(long long)b * do_bs + c]));
}

#define NUWA_CPL_SWITCH(cpl, CALL)      \
  switch (cpl) {                        \
    case 16: { CALL(16); } break;       \
    case 8: { CALL(8); } break;         \
    case 4: { CALL(4); } break;         \
    case 2: { CALL(2); } break;         \
    case 1: { CALL(1); } break;         \
    default: return NUWA_ERR_INVALID;   \
  }

template <int MODE>
static int launch_gather_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                                cudaStream_t stream) {
  const int inner = p.H * p.dh;
  if (inner % 32 != 0 || (32 % p.H) != 0 || p.H > 32 || jp < p.jmax) return NUWA_ERR_INVALID;
  const int cpl = inner / 32, warps = 4;
  const size_t smem = (size_t)warps * (2 * p.H * p.jmax + p.jmax) * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  const long long groups = (long long)p.B * p.nq;
  if (groups <= 0) return NUWA_OK;
  const unsigned grid = (unsigned)((groups + warps - 1) / warps);
#define CALL(C)                                                                                                      \
  if (smem > 48 * 1024)                                                                                              \
    cudaFuncSetAttribute(gather_scores_kernel<MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
  gather_scores_kernel<MODE, C><<<grid, warps * 32, smem, stream>>>(p, reinterpret_cast<const bf16*>(dO), do_bs, do_rs, S, dPp, jp)
  NUWA_CPL_SWITCH(cpl, CALL)
#undef CALL
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
template <int MODE>
static int launch_gather_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs,
                            cudaStream_t stream) {
  const int inner = p.H * p.dh;
  if (inner % 32 != 0 || (32 % p.H) != 0 || p.H > 32 || jp < p.jmax) return NUWA_ERR_INVALID;
  const int cpl = inner / 32, warps = 4;
  const size_t smem = (size_t)warps * (p.H * p.jmax + p.jmax) * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  const long long groups = (long long)p.B * p.nq;
  if (groups <= 0) return NUWA_OK;
  const unsigned grid = (unsigned)((groups + warps - 1) / warps);
#define CALL(C)                                                                                                  \
  if (smem > 48 * 1024)                                                                                          \
    cudaFuncSetAttribute(gather_dq_kernel<MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
  gather_dq_kernel<MODE, C><<<grid, warps * 32, smem, stream>>>(p, reinterpret_cast<const bf16*>(dS), jp,        \
                                                                 reinterpret_cast<bf16*>(dq), dq_bs, dq_rs)
  NUWA_CPL_SWITCH(cpl, CALL)
#undef CALL
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

int attn3dna_bwd_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                        cudaStream_t s) {
  return launch_gather_scores<MODE_3DNA>(p, dO, do_bs, do_rs, S, dPp, jp, s);
}
int attn3dna_bwd_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t s) {
  return launch_gather_dq<MODE_3DNA>(p, dS, jp, dq, dq_bs, dq_rs, s);
}
int attn3dna_bwd_dkdv(const AttnParams& p, const void* dO, long long do_bs, int do_rs, const void* dS, const void* Pp, int jp,
                      void* dk, void* dv, long long dkv_bs, int dkv_rs, cudaStream_t stream) {
  const int inner = p.H * p.dh;
  if (inner % 32 != 0 || (32 % p.H) != 0 || p.H > 32 || p.jmax > 255 || p.nv <= 0) return NUWA_ERR_INVALID;
  const int cpl = inner / 32, warps = 4;
  const size_t smem = (size_t)warps * p.jmax * sizeof(int);
  const long long groups = (long long)p.B * p.nv;
  const unsigned grid = (unsigned)((groups + warps - 1) / warps);
#define CALL(C)                                                                                                      \
  gather_dkdv_3dna_kernel<C><<<grid, warps * 32, smem, stream>>>(                                                    \
      p, reinterpret_cast<const bf16*>(dO), do_bs, do_rs, reinterpret_cast<const bf16*>(dS),                         \
      reinterpret_cast<const bf16*>(Pp), jp, reinterpret_cast<bf16*>(dk), reinterpret_cast<bf16*>(dv), dkv_bs, dkv_rs)
  NUWA_CPL_SWITCH(cpl, CALL)
#undef CALL
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
// SparseCross2DNA key-centric pass: context key (f_s, yy, xx) <- queries at (y, x) = key - offset*dil + pad in EVERY video
// frame (inverse of key_of<MODE_X2DNA>).  base_k / base_v (fp32, optional): the dense bos-query contribution, added in.
template <int CPL>
__global__ void __launch_bounds__(128)
gather_dkdv_x2dna_kernel(const AttnParams p, int nk, const bf16* __restrict__ dO, long long do_bs, int do_rs,
                         const bf16* __restrict__ dS, const bf16* __restrict__ Pp, int jp, const float* __restrict__ base_k,
                         const float* __restrict__ base_v, long long base_bs, int base_rs, bf16* __restrict__ dk,
                         bf16* __restrict__ dv, long long dkv_bs, int dkv_rs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int lph = 32 / p.H, h = lane / lph;
  const long long gidx = blockIdx.x * (long long)wpb + warp;
  if (gidx >= (long long)p.B * nk) return;
  const int b = (int)(gidx / nk), kidx = (int)(gidx - (long long)b * nk);
  const int T = p.fmap * p.fmap;
  const int fs = kidx / T, ky = (kidx % T) / p.fmap, kx = kidx % p.fmap;
  const int pad = p.cdil * (p.ck - 1) / 2;
  const int J2 = p.ck * p.ck;
  const int ch = lane * CPL;
  const bf16* qb = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  const bf16* dob = dO + (long long)b * do_bs;
  float ak[CPL], av[CPL];
  if (base_k != nullptr) {
    load_row_f32<CPL>(base_k + (long long)b * base_bs + (long long)kidx * base_rs + ch, ak);
    load_row_f32<CPL>(base_v + (long long)b * base_bs + (long long)kidx * base_rs + ch, av);
  } else {
#pragma unroll
    for (int c = 0; c < CPL; ++c) ak[c] = av[c] = 0.f;
  }
  const long long hb = ((long long)b * p.H + h) * p.nq;
  for (int w = 0; w < J2; ++w) {
    const int a = w / p.ck, c = w - a * p.ck;
    const int qy = ky - a * p.cdil + pad, qx = kx - c * p.cdil + pad;
    if (qy < 0 || qy >= p.fmap || qx < 0 || qx >= p.fmap) continue;
    const int j = 1 + fs * J2 + w;
    for (int ql = qy * p.fmap + qx + 1 - p.t0; ql < p.nq; ql += T) {  // the same (y, x) in every video frame
      if (ql < 0) continue;
      const float ws = __bfloat162float(dS[(hb + ql) * jp + j]);
      const float wp = __bfloat162float(Pp[(hb + ql) * jp + j]);
      float qv[CPL], dv_[CPL];
      load_row<CPL>(qb + (long long)ql * p.q_rs + ch, qv);
      load_row<CPL>(dob + (long long)ql * do_rs + ch, dv_);
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) {
        ak[cc] = fmaf(ws, qv[cc], ak[cc]);
        av[cc] = fmaf(wp, dv_[cc], av[cc]);
      }
    }
  }
  const long long o = (long long)b * dkv_bs + (long long)kidx * dkv_rs + ch;
  store_row<CPL>(dk + o, ak);
  store_row<CPL>(dv + o, av);
}

int attnx2_bwd_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                      cudaStream_t s) {
  return launch_gather_scores<MODE_X2DNA>(p, dO, do_bs, do_rs, S, dPp, jp, s);
}
int attnx2_bwd_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t s) {
  return launch_gather_dq<MODE_X2DNA>(p, dS, jp, dq, dq_bs, dq_rs, s);
}
int attnx2_bwd_dkdv(const AttnParams& p, int nk, const void* dO, long long do_bs, int do_rs, const void* dS, const void* Pp,
                    int jp, const float* base_k, const float* base_v, long long base_bs, int base_rs, void* dk, void* dv,
                    long long dkv_bs, int dkv_rs, cudaStream_t stream) {
  const int inner = p.H * p.dh;
  if (inner % 32 != 0 || (32 % p.H) != 0 || p.H > 32 || nk <= 0 || p.fmap <= 0 || (nk % (p.fmap * p.fmap)) != 0)
    return NUWA_ERR_INVALID;
  const int cpl = inner / 32, warps = 4;
  const long long groups = (long long)p.B * nk;
  const unsigned grid = (unsigned)((groups + warps - 1) / warps);
#define CALL(C)                                                                                                         \
  gather_dkdv_x2dna_kernel<C><<<grid, warps * 32, 0, stream>>>(                                                         \
      p, nk, reinterpret_cast<const bf16*>(dO), do_bs, do_rs, reinterpret_cast<const bf16*>(dS),                        \
      reinterpret_cast<const bf16*>(Pp), jp, base_k, base_v, base_bs, base_rs, reinterpret_cast<bf16*>(dk),             \
      reinterpret_cast<bf16*>(dv), dkv_bs, dkv_rs)
  NUWA_CPL_SWITCH(cpl, CALL)
#undef CALL
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
int attn_bwd_first_key(const void* q, long long q_bs, int q_rs, const void* dO, long long do_bs, int do_rs, const void* dS,
                       const void* Pp, int jp, int B, int H, int dh, int nq, float* out_k, float* out_v, long long ok_bs,
                       cudaStream_t stream) {
  const int inner = H * dh;
  if (inner > 1024 || inner <= 0 || B <= 0 || nq <= 0) return NUWA_ERR_INVALID;
  const int chunk = 64;
  dim3 grid(ceil_div(nq, chunk), B);
  first_key_kernel<<<grid, inner, 2 * H * chunk * sizeof(float), stream>>>(reinterpret_cast<const bf16*>(q), q_bs, q_rs, reinterpret_cast<const bf16*>(dO),
                                               do_bs, do_rs, reinterpret_cast<const bf16*>(dS),
                                               reinterpret_cast<const bf16*>(Pp), jp, H, dh, nq, chunk, out_k, out_v, ok_bs);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
int attn3dna_bwd_first_key_finalize(const float* tmp_k, const float* tmp_v, const void* dO_bos, long long do_bs, void* dqkv,
                                    long long dqkv_bs, int inner, int B, cudaStream_t stream) {
  first_key_finalize_kernel<<<ceil_div(B * inner, 256), 256, 0, stream>>>(tmp_k, tmp_v, reinterpret_cast<const bf16*>(dO_bos),
                                                                          do_bs, reinterpret_cast<bf16*>(dqkv), dqkv_bs, inner, B);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
