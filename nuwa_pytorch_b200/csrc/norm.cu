// Row-wise normalisation kernels of the transformer stacks (HBM-bound, one warp per token row):
//
//  sandwich_ln:  [stage A]  x_out = res_in + LayerNorm_post(y)          (SandwichNorm tail + residual,
//                                                                        nuwa_pytorch.py:127,1175-1180)
//                [stage B]  a = LayerNorm_pre(x_out)  -> bf16 GEMM operand, with the ShiftVideoTokens
//                           channel shift fused as a SCATTER (nuwa_pytorch.py:200-253): the first D/4
//                           channels of token t are written to the row of the token one grid-row below,
//                           the second D/4 to the token one column right; border rows get zeros.
//                           Scatter form makes the op valid for incremental decode too (the shifted
//                           channels simply wait in the persistent operand buffer for their token).
//  stable_ln:    StableLayerNorm of (a [+ b]) : LN(v / amax(v))          (nuwa_pytorch.py:88-95, REV:142)
//
// The same stage-A kernel implements LayerNormChan + residual of VQGanAttention on NHWC rows
// (vqgan_vae.py:140-143,286).
#include "common.cuh"
#include "kernels.h"

namespace nuwa {

// MAXV = float4 values per lane: 8 -> D <= 1024 (transformer rows), 32 -> D <= 4096 (VAE channel rows)

template <int LN_MAXV>
__device__ __forceinline__ void row_stats(const float4 (&v)[LN_MAXV], int nv, int D, float& mean, float& rstd,
                                          float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
  s = warp_sum(s);
  mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  q = warp_sum(q);
  rstd = rsqrtf(q / (float)D + eps);
}

template <int LN_MAXV>
__global__ void __launch_bounds__(256) sandwich_ln_kernel(const LnParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = p.B * p.nt;
  if (warp >= rows) return;
  const int b = warp / p.nt;
  const int tl = warp - b * p.nt;
  const int t0 = p.t0_ptr != nullptr ? __ldg(p.t0_ptr) : p.t0;
  const int t = t0 + tl;
  const int D = p.D;
  // number of float4 this lane owns: channels c = (lane + 32*i)*4
  int nv = 0;
  float4 v[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) nv = i + 1;
  }
  const long long roff = (long long)warp * D;
  if (p.y != nullptr) {
    // ---- stage A : x_out = res_in + LN_post(y) ----
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) v[i] = *reinterpret_cast<const float4*>(p.y + roff + (lane + 32 * i) * 4);
    float mean, rstd;
    row_stats<LN_MAXV>(v, nv, D, mean, rstd, p.eps);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        const float4 w = *reinterpret_cast<const float4*>(p.post_w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.post_b + c);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.res_in != nullptr) r = *reinterpret_cast<const float4*>(p.res_in + roff + c);
        v[i].x = (v[i].x - mean) * rstd * w.x + bb.x + r.x;
        v[i].y = (v[i].y - mean) * rstd * w.y + bb.y + r.y;
        v[i].z = (v[i].z - mean) * rstd * w.z + bb.z + r.z;
        v[i].w = (v[i].w - mean) * rstd * w.w + bb.w + r.w;
        if (p.x_out != nullptr) *reinterpret_cast<float4*>(p.x_out + roff + c) = v[i];
        if (p.x_out_bf16 != nullptr) {
          uint2 pk;
          pk.x = pack_bf16x2(v[i].x, v[i].y);
          pk.y = pack_bf16x2(v[i].z, v[i].w);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.x_out_bf16) + roff + c) = pk;
        }
      }
  } else {
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) v[i] = *reinterpret_cast<const float4*>(p.res_in + roff + (lane + 32 * i) * 4);
  }
  if (p.pre_w == nullptr) return;
  // ---- stage B : a = LN_pre(x) -> bf16, optional shift scatter ----
  float mean, rstd;
  row_stats<LN_MAXV>(v, nv, D, mean, rstd, p.eps);
  const int q4 = D / 4;
  int dst_h = -1, dst_w = -1;  // destination rows (absolute positions) of the two shifted chunks
  bool zero_h = false, zero_w = false;
  if (p.shift && t >= 1) {
    const int T = p.fmap * p.fmap;
    const int pos = (t - 1) % T;
    const int row = pos / p.fmap, col = pos - row * p.fmap;
    zero_h = (row == 0);
    zero_w = (col == 0);
    if (row < p.fmap - 1 && t + p.fmap < p.a_npos) dst_h = t + p.fmap;
    if (col < p.fmap - 1 && t + 1 < p.a_npos) dst_w = t + 1;
  }
  bf16* abase = reinterpret_cast<bf16*>(p.a_out) + (long long)b * p.a_bs;
  if (p.gather) {
    // decode form (one CUDA graph replayed per token): fixed-address dense operand row; the shifted channels of
    // position t were produced when positions t - fmap / t - 1 were decoded and wait in shift_cache.
    bf16* sc = reinterpret_cast<bf16*>(p.shift_cache) + (long long)b * p.sc_bs;
    const bool shifted = p.shift && t >= 1;
    int src_h = -1, src_w = -1;
    if (shifted) {
      const int T = p.fmap * p.fmap;
      const int pos = (t - 1) % T;
      const int row = pos / p.fmap, col = pos - row * p.fmap;
      if (row > 0) src_h = t - p.fmap;
      if (col > 0) src_w = t - 1;
    }
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        const float4 w = *reinterpret_cast<const float4*>(p.pre_w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.pre_b + c);
        uint2 pk;
        pk.x = pack_bf16x2((v[i].x - mean) * rstd * w.x + bb.x, (v[i].y - mean) * rstd * w.y + bb.y);
        pk.y = pack_bf16x2((v[i].z - mean) * rstd * w.z + bb.z, (v[i].w - mean) * rstd * w.w + bb.w);
        uint2 outv = pk;
        if (p.shift && c < 2 * q4) {
          if (t < p.a_npos) *reinterpret_cast<uint2*>(sc + (long long)t * D + c) = pk;  // for the tokens to come
          if (shifted) {
            const int src = c < q4 ? src_h : src_w;
            outv = src >= 0 ? *reinterpret_cast<const uint2*>(sc + (long long)src * D + c) : make_uint2(0u, 0u);
          }
        }
        *reinterpret_cast<uint2*>(abase + (long long)tl * p.a_rs + c) = outv;
      }
    return;
  }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const int c = (lane + 32 * i) * 4;
      const float4 w = *reinterpret_cast<const float4*>(p.pre_w + c);
      const float4 bb = *reinterpret_cast<const float4*>(p.pre_b + c);
      uint2 pk;
      pk.x = pack_bf16x2((v[i].x - mean) * rstd * w.x + bb.x, (v[i].y - mean) * rstd * w.y + bb.y);
      pk.y = pack_bf16x2((v[i].z - mean) * rstd * w.z + bb.z, (v[i].w - mean) * rstd * w.w + bb.w);
      const uint2 zero = make_uint2(0u, 0u);
      bf16* own = abase + (long long)(t - p.a_t0) * p.a_rs + c;
      if (!p.shift || t == 0 || c >= 2 * q4) {
        *reinterpret_cast<uint2*>(own) = pk;
      } else if (c < q4) {
        if (dst_h >= 0) *reinterpret_cast<uint2*>(abase + (long long)(dst_h - p.a_t0) * p.a_rs + c) = pk;
        if (zero_h) *reinterpret_cast<uint2*>(own) = zero;
      } else {
        if (dst_w >= 0) *reinterpret_cast<uint2*>(abase + (long long)(dst_w - p.a_t0) * p.a_rs + c) = pk;
        if (zero_w) *reinterpret_cast<uint2*>(own) = zero;
      }
    }
}

int sandwich_ln(const LnParams& p, cudaStream_t stream) {
  if (p.D % 16 != 0 || p.D > 32 * 128 || p.B <= 0 || p.nt <= 0) return NUWA_ERR_INVALID;
  if (p.y == nullptr && p.res_in == nullptr) return NUWA_ERR_INVALID;
  if (p.y != nullptr && (p.post_w == nullptr || p.post_b == nullptr)) return NUWA_ERR_INVALID;
  if (p.pre_w != nullptr && (p.pre_b == nullptr || p.a_out == nullptr)) return NUWA_ERR_INVALID;
  const int rows = p.B * p.nt;
  const int wpb = 8;
  if (p.D <= 1024) sandwich_ln_kernel<8><<<ceil_div(rows, wpb), wpb * 32, 0, stream>>>(p);
  else sandwich_ln_kernel<32><<<ceil_div(rows, wpb), wpb * 32, 0, stream>>>(p);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

template <int LN_MAXV>
__global__ void __launch_bounds__(256)
stable_ln_kernel(const float* __restrict__ a, const float* __restrict__ b2, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int rows,
                 int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long roff = (long long)warp * D;
  float4 v[LN_MAXV];
  int nv = 0;
  float mx = -3.402823466e38f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      nv = i + 1;
      v[i] = *reinterpret_cast<const float4*>(a + roff + c);
      if (b2 != nullptr) {
        const float4 u = *reinterpret_cast<const float4*>(b2 + roff + c);
        v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
      }
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  mx = warp_max(mx);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      v[i].x /= mx; v[i].y /= mx; v[i].z /= mx; v[i].w /= mx;  // x / amax(x)  (true division, as the reference)
    }
  float mean, rstd;
  row_stats<LN_MAXV>(v, nv, D, mean, rstd, eps);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const int c = (lane + 32 * i) * 4;
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      const float4 bb = *reinterpret_cast<const float4*>(bias + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
      o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
      o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
      o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + roff + c) = o;
      if (out_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(out_bf16 + roff + c) = pk;
      }
    }
}

int stable_ln(const float* a, const float* b2, const float* w, const float* bias, float* out_f32, void* out_bf16,
              int rows, int D, cudaStream_t stream) {
  if (D % 16 != 0 || D > 8 * 128 || rows <= 0) return NUWA_ERR_INVALID;
  const int wpb = 8;
  stable_ln_kernel<8><<<ceil_div(rows, wpb), wpb * 32, 0, stream>>>(a, b2, w, bias, out_f32,
                                                                  reinterpret_cast<bf16*>(out_bf16), rows, D, 1e-5f);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
