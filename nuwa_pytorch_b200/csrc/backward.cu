// Backward (training) kernels of the transformer stacks -- the gradient of the path timed by BASELINE configs[2]/[4]
// (`loss = nuwa(text, video, return_loss=True); loss.backward()`, nuwa_pytorch.py:1917-1964 differentiated).
//
//  ln_bwd            LayerNorm / StableLayerNorm backward of one row per warp (SandwichNorm pre/post norms,
//                    nuwa_pytorch.py:112-128, 88-95), with the inverse ShiftVideoTokens map (:200-253) fused as a gather
//                    and per-CTA partial sums of dweight / dbias / column-sum(dx) (the latter is the bias gradient of the
//                    linear layer that produced the normalised tensor).
//  reduce_partials   out[c] += sum_p part[p][c]
//  transpose_bf16    [R][C] -> [C][R] (weight-gradient GEMM operands: the contraction index must be contiguous for TMA)
//  geglu_fwd/bwd     GEGLU on the pair-packed pre-activation (nuwa_pytorch.py:255-258)
//  ce_bwd            d(mean cross entropy)/d(logits) -> bf16 (nuwa_pytorch.py:1963)
//  embed_bwd         scatter-add into embedding table (x frac_gradient, :1666-1670), axial position tables and bos
//  rotary_bwd        inverse rotation of dq,dk,dv (:132-153)
//  add_rows_f32      dst[map[r]] (+)= src[r]  (un-packs pair-packed / padded weight gradients)
//  attention:        dense  -> kv_full_build, attn_bwd_rows, kv_full_split   (+ bgemm.cu for the five products)
//                    gather -> gather_scores, attn_bwd_rows, gather_dq, gather_dkdv, gather_first_key
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

// =================================================================================================
// LayerNorm backward
// =================================================================================================
template <int MAXV, bool COLSUM>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const nuwa_lnbwd_params p) {
  extern __shared__ float sm_part[];  // [3][D] CTA partial sums
  const int D = p.D;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sm_part[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long nwarps = (long long)gridDim.x * wpb;
  int nv = 0;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if ((lane + 32 * i) * 4 < D) nv = i + 1;
  float4 aw[MAXV], ab[MAXV], ac[COLSUM ? MAXV : 1];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) aw[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (COLSUM ? MAXV : 1); ++i) ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int q4 = D / 4;
  const float invD = 1.0f / (float)D;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < p.rows; row += nwarps) {
    const int b = (int)(row / p.nt), t = (int)(row - (long long)b * p.nt);
    const long long roff = row * D;
    // ---- upstream gradient rows (through the inverse token shift when the normalised row fed a shifted sub-block) ----
    int src_h = t, src_w = t;  // rows the first / second channel quarter of the normalised row went to
    bool ok_h = true, ok_w = true;
    if (p.unshift && t >= 1) {
      const int T = p.fmap * p.fmap;
      const int pos = (t - 1) % T;
      const int gr = pos / p.fmap, gc = pos - gr * p.fmap;
      ok_h = (gr < p.fmap - 1) && (t + p.fmap < p.nt);
      ok_w = (gc < p.fmap - 1) && (t + 1 < p.nt);
      src_h = t + p.fmap;
      src_w = t + 1;
    }
    // ---- all loads of the row first (memory-level parallelism), then the reductions ----
    float4 v[MAXV], g[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        v[i] = *reinterpret_cast<const float4*>(p.x + roff + c);
        int sr = t;
        bool ok = true;
        if (p.unshift && t >= 1 && c < 2 * q4) {
          sr = c < q4 ? src_h : src_w;
          ok = c < q4 ? ok_h : ok_w;
        }
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          const long long so = ((long long)b * p.nt + sr) * D + c;
          if (p.dout_f32 != nullptr) d = *reinterpret_cast<const float4*>(p.dout_f32 + so);
          else {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.dout_bf16) + so);
            const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
            d = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
        }
        g[i] = d;
      }
    if (p.x2 != nullptr) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
          const float4 u = *reinterpret_cast<const float4*>(p.x2 + roff + (lane + 32 * i) * 4);
          v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
        }
    }
    float inv_mx = 1.0f;
    if (p.stable) {
      float mx = -FLT_MAX;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (i < nv) mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      mx = warp_max(mx);
      inv_mx = 1.0f / mx;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (i < nv) { v[i].x /= mx; v[i].y /= mx; v[i].z /= mx; v[i].w /= mx; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(s) * invD;
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        qv += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    const float rstd = rsqrtf(warp_sum(qv) * invD + p.eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w + (lane + 32 * i) * 4));
        const float4 d = g[i];
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
        aw[i].x += d.x * v[i].x; aw[i].y += d.y * v[i].y; aw[i].z += d.z * v[i].z; aw[i].w += d.w * v[i].w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        g[i].x = d.x * w4.x; g[i].y = d.y * w4.y; g[i].z = d.z * w4.z; g[i].w = d.w * w4.w;  // dout * w
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * v[i].x + g[i].y * v[i].y + g[i].z * v[i].z + g[i].w * v[i].w;
      }
    const float m1 = warp_sum(s1) * invD, m2 = warp_sum(s2) * invD;
    const float k = rstd * inv_mx;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        float4 dx;
        dx.x = k * (g[i].x - m1 - v[i].x * m2);
        dx.y = k * (g[i].y - m1 - v[i].y * m2);
        dx.z = k * (g[i].z - m1 - v[i].z * m2);
        dx.w = k * (g[i].w - m1 - v[i].w * m2);
        if (COLSUM) { ac[i].x += dx.x; ac[i].y += dx.y; ac[i].z += dx.z; ac[i].w += dx.w; }
        if (p.dx_bf16 != nullptr) {
          uint2 pk;
          pk.x = pack_bf16x2(dx.x, dx.y);
          pk.y = pack_bf16x2(dx.z, dx.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.dx_bf16) + roff + c) = pk;
        }
        if (p.dx_f32 != nullptr) {
          float4* dst = reinterpret_cast<float4*>(p.dx_f32 + roff + c);
          float4 o = dx;
          if (p.accumulate) { const float4 e = *dst; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
          *dst = o;
        }
        if (p.dx2_f32 != nullptr) {
          float4* dst = reinterpret_cast<float4*>(p.dx2_f32 + roff + c);
          float4 o = dx;
          if (p.accumulate) { const float4 e = *dst; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
          *dst = o;
        }
      }
  }
  // ---- parameter gradients: CTA reduction in shared memory, then one fp32 atomic per (CTA, channel) ----
  if (p.dw != nullptr || p.db != nullptr || p.dcol != nullptr) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 4;
        atomicAdd(&sm_part[c + 0], aw[i].x); atomicAdd(&sm_part[c + 1], aw[i].y);
        atomicAdd(&sm_part[c + 2], aw[i].z); atomicAdd(&sm_part[c + 3], aw[i].w);
        atomicAdd(&sm_part[D + c + 0], ab[i].x); atomicAdd(&sm_part[D + c + 1], ab[i].y);
        atomicAdd(&sm_part[D + c + 2], ab[i].z); atomicAdd(&sm_part[D + c + 3], ab[i].w);
        if (COLSUM) {
          atomicAdd(&sm_part[2 * D + c + 0], ac[i].x); atomicAdd(&sm_part[2 * D + c + 1], ac[i].y);
          atomicAdd(&sm_part[2 * D + c + 2], ac[i].z); atomicAdd(&sm_part[2 * D + c + 3], ac[i].w);
        }
      }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      if (p.dw != nullptr) atomicAdd(p.dw + i, sm_part[i]);
      if (p.db != nullptr) atomicAdd(p.db + i, sm_part[D + i]);
      if (COLSUM && p.dcol != nullptr) atomicAdd(p.dcol + i, sm_part[2 * D + i]);
    }
  }
}

int ln_bwd_grid(int rows) {
  const int want = ceil_div(rows, 8 * 4);  // >= 4 rows per warp amortise the parameter-gradient reduction
  const int cap = device_sm_count() * 2;
  return want < 1 ? 1 : (want > cap ? cap : want);
}

int ln_bwd(const nuwa_lnbwd_params& p, cudaStream_t stream) {
  if (p.D % 16 != 0 || p.D > 1024 || p.rows <= 0 || p.nt <= 0 || (p.rows % p.nt) != 0) return NUWA_ERR_INVALID;
  if ((p.dout_f32 == nullptr) == (p.dout_bf16 == nullptr)) return NUWA_ERR_INVALID;
  if (p.x == nullptr || p.w == nullptr) return NUWA_ERR_INVALID;
  if (p.unshift && p.fmap <= 0) return NUWA_ERR_INVALID;
  const int grid = ln_bwd_grid(p.rows);
  const size_t smem = (size_t)3 * p.D * sizeof(float);
  const bool cs = p.dcol != nullptr;
  if (p.D <= 512) {
    if (cs) ln_bwd_kernel<4, true><<<grid, 256, smem, stream>>>(p);
    else ln_bwd_kernel<4, false><<<grid, 256, smem, stream>>>(p);
  } else {
    if (cs) ln_bwd_kernel<8, true><<<grid, 256, smem, stream>>>(p);
    else ln_bwd_kernel<8, false><<<grid, 256, smem, stream>>>(p);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// out_k[c] += sum_p part[p][k][c]   (k = 0: dweight, 1: dbias, 2: column sum of dx) ; NULL outputs are skipped
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int nparts, int D, float* __restrict__ o0, float* __restrict__ o1,
                       float* __restrict__ o2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * D) return;
  const int k = i / D, c = i - k * D;
  float* o = k == 0 ? o0 : (k == 1 ? o1 : o2);
  if (o == nullptr) return;
  float s = 0.f;
  for (int pp = 0; pp < nparts; ++pp) s += part[(long long)pp * 3 * D + i];
  o[c] += s;
}
int reduce_partials(const float* part, int nparts, int D, float* o0, float* o1, float* o2, cudaStream_t stream) {
  if (nparts <= 0 || D <= 0) return NUWA_ERR_INVALID;
  reduce_partials_kernel<<<ceil_div(3 * D, 256), 256, 0, stream>>>(part, nparts, D, o0, o1, o2);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// bf16 transpose: in [R][ld_in] (C valid columns) -> out [C][ld_out] (R valid columns)
// =================================================================================================
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const bf16* __restrict__ in, long long ld_in, bf16* __restrict__ out, long long ld_out, int R, int C) {
  __shared__ __align__(16) unsigned short tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = (tid >> 3) + 32 * pass, cc = (tid & 7) * 8;
    const int gr = r0 + r, gc = c0 + cc;
    unsigned short e[8];
    if (gr < R && gc + 8 <= C && ((ld_in & 7) == 0)) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + (long long)gr * ld_in + gc));
      e[0] = u.x & 0xffff; e[1] = u.x >> 16; e[2] = u.y & 0xffff; e[3] = u.y >> 16;
      e[4] = u.z & 0xffff; e[5] = u.z >> 16; e[6] = u.w & 0xffff; e[7] = u.w >> 16;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        e[i] = (gr < R && gc + i < C) ? reinterpret_cast<const unsigned short*>(in)[(long long)gr * ld_in + gc + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[r][cc + i] = e[i];
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int c = (tid >> 3) + 32 * pass, rr = (tid & 7) * 8;
    const int gc = c0 + c, gr = r0 + rr;
    if (gc >= C || gr >= R) continue;
    unsigned short e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = tile[rr + i][c];
    unsigned short* dst = reinterpret_cast<unsigned short*>(out) + (long long)gc * ld_out + gr;
    if (gr + 8 <= R && ((ld_out & 7) == 0)) {
      uint4 u;
      u.x = e[0] | ((uint32_t)e[1] << 16); u.y = e[2] | ((uint32_t)e[3] << 16);
      u.z = e[4] | ((uint32_t)e[5] << 16); u.w = e[6] | ((uint32_t)e[7] << 16);
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      for (int i = 0; i < 8 && gr + i < R; ++i) dst[i] = e[i];
    }
  }
}
// Same transpose with 32-bit shared-memory traffic: the tile is stored as bf16 PAIRS (two adjacent input columns per
// word); a thread reads the 8 words (rows rr..rr+7, column pair cp) and splits them with byte permutes into the two
// 16-byte output rows 2cp and 2cp+1 -- one pass, 8 LDS.32 per 16 output elements instead of 16 LDS.U16, and a warp's
// stores cover 128 contiguous bytes of 4 output rows.  Needs ld_in % 8 == 0, ld_out % 8 == 0 (16-byte vectors).
__global__ void __launch_bounds__(256)
transpose_bf16_pairs_kernel(const bf16* __restrict__ in, long long ld_in, bf16* __restrict__ out, long long ld_out, int R, int C) {
  __shared__ uint32_t tile[64][33];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = (tid >> 3) + 32 * pass, cc = (tid & 7) * 8;
    const int gr = r0 + r, gc = c0 + cc;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (gr < R && gc + 8 <= C) {
      u = __ldg(reinterpret_cast<const uint4*>(in + (long long)gr * ld_in + gc));
    } else if (gr < R && gc < C) {
      unsigned short e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = gc + i < C ? reinterpret_cast<const unsigned short*>(in)[(long long)gr * ld_in + gc + i] : 0;
      u.x = e[0] | ((uint32_t)e[1] << 16); u.y = e[2] | ((uint32_t)e[3] << 16);
      u.z = e[4] | ((uint32_t)e[5] << 16); u.w = e[6] | ((uint32_t)e[7] << 16);
    }
    uint32_t* t = &tile[r][cc >> 1];
    t[0] = u.x; t[1] = u.y; t[2] = u.z; t[3] = u.w;
  }
  __syncthreads();
  const int cp = tid >> 3, rr = (tid & 7) * 8;  // column pair, first of 8 rows
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = tile[rr + i][cp];
  const int gr = r0 + rr;
  if (gr >= R) return;
#pragma unroll
  for (int hsel = 0; hsel < 2; ++hsel) {
    const int gc = c0 + 2 * cp + hsel;
    if (gc >= C) continue;
    const uint32_t sel = hsel ? 0x7632u : 0x5410u;  // high / low halves of (a, b)
    uint4 o;
    o.x = __byte_perm(w[0], w[1], sel); o.y = __byte_perm(w[2], w[3], sel);
    o.z = __byte_perm(w[4], w[5], sel); o.w = __byte_perm(w[6], w[7], sel);
    unsigned short* dst = reinterpret_cast<unsigned short*>(out) + (long long)gc * ld_out + gr;
    if (gr + 8 <= R) {
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
      for (int i = 0; i < 8 && gr + i < R; ++i) dst[i] = (unsigned short)(ow[i >> 1] >> (16 * (i & 1)));
    }
  }
}
int transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int R, int C, cudaStream_t stream) {
  if (R <= 0 || C <= 0 || ld_in < C || ld_out < R) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return NUWA_ERR_INVALID;
  dim3 grid(ceil_div(C, 64), ceil_div(R, 64));
  if (grid.y > 65535) return NUWA_ERR_INVALID;
  if ((ld_in & 7) == 0 && (ld_out & 7) == 0) {
    transpose_bf16_pairs_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), ld_in, reinterpret_cast<bf16*>(out),
                                                          ld_out, R, C);
    NUWA_CHECK_LAUNCH();
    return NUWA_OK;
  }
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), ld_in, reinterpret_cast<bf16*>(out),
                                                  ld_out, R, C);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// GEGLU on the pair-packed pre-activation: h [M][2*ip], every 32 columns = 16 values then their 16 gates
// =================================================================================================
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ g, long long M, int ip) {
  const int cpr = ip / 8;  // 8-wide output chunks per row
  const long long total = M * cpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / cpr;
    const int o = (int)(i - m * cpr) * 8;  // output column
    const int src = (o >> 4) * 32 + (o & 15);
    float a[8], gt[8], r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + m * 2 * ip + src)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + m * 2 * ip + src + 16)), gt);
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = a[e] * gelu_erf(gt[e]);
    *reinterpret_cast<uint4*>(g + m * ip + o) = pack8(r);
  }
}
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const bf16* __restrict__ dg, const bf16* __restrict__ h, bf16* __restrict__ dh, long long M, int ip) {
  const int cpr = ip / 8;
  const long long total = M * cpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / cpr;
    const int o = (int)(i - m * cpr) * 8;
    const int src = (o >> 4) * 32 + (o & 15);
    float a[8], gt[8], d[8], da[8], dgt[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + m * 2 * ip + src)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + m * 2 * ip + src + 16)), gt);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dg + m * ip + o)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = gt[e];
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
      da[e] = d[e] * x * cdf;
      dgt[e] = d[e] * a[e] * (cdf + x * pdf);
    }
    *reinterpret_cast<uint4*>(dh + m * 2 * ip + src) = pack8(da);
    *reinterpret_cast<uint4*>(dh + m * 2 * ip + src + 16) = pack8(dgt);
  }
}
static int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
int geglu_fwd(const void* h, void* g, long long M, int ip, cudaStream_t stream) {
  if (M <= 0 || ip <= 0 || (ip % 16)) return NUWA_ERR_INVALID;
  geglu_fwd_kernel<<<ew_grid(M * (ip / 8)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(h), reinterpret_cast<bf16*>(g), M, ip);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
int geglu_bwd(const void* dg, const void* h, void* dh, long long M, int ip, cudaStream_t stream) {
  if (M <= 0 || ip <= 0 || (ip % 16)) return NUWA_ERR_INVALID;
  geglu_bwd_kernel<<<ew_grid(M * (ip / 8)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dg), reinterpret_cast<const bf16*>(h),
                                                            reinterpret_cast<bf16*>(dh), M, ip);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// cross entropy backward: dlogits[r][c] = (softmax(logits[r])[c] - [c == target[r]]) * (*gscale) / rows   -> bf16
// =================================================================================================
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ target, const float* __restrict__ gscale,
              bf16* __restrict__ dl, int ld_out, int rows, int V) {
  const int row = blockIdx.x;
  const float* l = logits + (long long)row * ld;
  __shared__ float red[32];
  __shared__ float bc;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -FLT_MAX;
  for (int c = threadIdx.x; c < V; c += blockDim.x) m = fmaxf(m, l[c]);
  m = warp_max(m);
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int i = 1; i < nw; ++i) mm = fmaxf(mm, red[i]);
    bc = mm;
  }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) s += __expf(l[c] - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f;
    for (int i = 0; i < nw; ++i) ss += red[i];
    bc = ss;
  }
  __syncthreads();
  const float k = (gscale != nullptr ? __ldg(gscale) : 1.0f) / (float)rows;
  const float inv = k / bc;
  const int tgt = (int)target[row];
  bf16* o = dl + (long long)row * ld_out;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float v = __expf(l[c] - m) * inv;
    if (c == tgt) v -= k;
    o[c] = __float2bfloat16(v);
  }
}
// Same with the row held in registers (V <= 256 * 4 * NV): one 16-byte-vectorised pass over the logits, 8-byte stores.
template <int NV>
__global__ void __launch_bounds__(256)
ce_bwd_reg_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ target,
                  const float* __restrict__ gscale, bf16* __restrict__ dl, int ld_out, int rows, int V) {
  const int row = blockIdx.x;
  const float4* l4 = reinterpret_cast<const float4*>(logits + (long long)row * ld);
  const int nv4 = V >> 2;
  __shared__ float red[2][8];
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = c < nv4 ? __ldcs(l4 + c) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
  float m = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[0][i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x = __expf(v[i].x - m); v[i].y = __expf(v[i].y - m); v[i].z = __expf(v[i].z - m); v[i].w = __expf(v[i].w - m);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[1][i];
  const float k = (gscale != nullptr ? __ldg(gscale) : 1.0f) / (float)rows;
  const float inv = k / tot;
  const int tgt = (int)target[row];
  uint2* o = reinterpret_cast<uint2*>(dl + (long long)row * ld_out);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c >= nv4) continue;
    float a = v[i].x * inv, b = v[i].y * inv, cc = v[i].z * inv, d = v[i].w * inv;
    if ((tgt >> 2) == c) {
      const int e = tgt & 3;
      if (e == 0) a -= k; else if (e == 1) b -= k; else if (e == 2) cc -= k; else d -= k;
    }
    uint2 u;
    u.x = pack_bf16x2(a, b);
    u.y = pack_bf16x2(cc, d);
    o[c] = u;
  }
}
int ce_bwd(const float* logits, int ld, const long long* target, const float* gscale, void* dlogits, int ld_out, int rows,
           int V, cudaStream_t stream) {
  if (rows <= 0 || V <= 0 || ld < V || ld_out < V) return NUWA_ERR_INVALID;
  const bool vec = (V % 4 == 0) && (ld % 4 == 0) && (ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dlogits) & 7) == 0);
  if (vec && V <= 256 * 4 * 8) {
    if (V <= 256 * 4 * 2)
      ce_bwd_reg_kernel<2><<<rows, 256, 0, stream>>>(logits, ld, target, gscale, reinterpret_cast<bf16*>(dlogits), ld_out, rows, V);
    else
      ce_bwd_reg_kernel<8><<<rows, 256, 0, stream>>>(logits, ld, target, gscale, reinterpret_cast<bf16*>(dlogits), ld_out, rows, V);
    NUWA_CHECK_LAUNCH();
    return NUWA_OK;
  }
  ce_bwd_kernel<<<rows, 256, 0, stream>>>(logits, ld, target, gscale, reinterpret_cast<bf16*>(dlogits), ld_out, rows, V);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// embedding backward (scatter-add; rows that share a destination are few enough for L2 atomics)
//   dx [B*nt][D]; position t of sample b:  bos (has_bos && t == 0) -> dbos ; else token idx[b][t - has_bos]:
//   dtable[token] += frac * dx ; dax1[p/(d2*d3)] += dx ; dax2[(p/d3)%d2] += dx ; dax3[p%d3] += dx
// =================================================================================================
__global__ void __launch_bounds__(256) embed_bwd_kernel(const nuwa_embed_bwd_params p) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)p.B * p.nt) return;
  const int b = (int)(row / p.nt), t = (int)(row - (long long)b * p.nt);
  const float* dx = p.dx + row * p.D;
  if (p.has_bos && t == 0) {
    if (p.dbos != nullptr)
      for (int c = lane; c < p.D; c += 32) atomicAdd(p.dbos + c, dx[c]);
    return;
  }
  const int pidx = t - (p.has_bos ? 1 : 0);
  const long long tok = p.idx[(long long)b * p.idx_bs + pidx];
  float* dt = p.dtable + tok * p.D;
  float* a1 = p.dax1 ? p.dax1 + (long long)(pidx / (p.d2 * p.d3)) * p.D : nullptr;
  float* a2 = p.dax2 ? p.dax2 + (long long)((pidx / p.d3) % p.d2) * p.D : nullptr;
  float* a3 = p.dax3 ? p.dax3 + (long long)(pidx % p.d3) * p.D : nullptr;
  for (int c = lane; c < p.D; c += 32) {
    const float g = dx[c];
    atomicAdd(dt + c, g * p.frac);
    if (a1) atomicAdd(a1 + c, g);
    if (a2) atomicAdd(a2 + c, g);
    if (a3) atomicAdd(a3 + c, g);
  }
}
int embed_bwd(const nuwa_embed_bwd_params& p, cudaStream_t stream) {
  if (p.B <= 0 || p.nt <= 0 || p.D <= 0 || p.dx == nullptr || p.dtable == nullptr || p.idx == nullptr) return NUWA_ERR_INVALID;
  const long long rows = (long long)p.B * p.nt;
  embed_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(p);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// rotary backward: dqkv (post-rotation gradient, fp32 [rows][3*H*dh]) -> pre-rotation gradient, bf16
// =================================================================================================
__global__ void __launch_bounds__(256)
rotary_bwd_kernel(const float* __restrict__ dqkv, bf16* __restrict__ out, const float* __restrict__ inv_freq, int rows, int n,
                  int H, int dh, int rot) {
  const long long total = (long long)rows * 3 * H * dh;
  const int r2 = rot / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dh);
    const long long row = i / (3LL * H * dh);
    const int t = (int)(row % n);
    float x = dqkv[i];
    if (d < rot) {
      // forward: y_lo = x_lo c - x_hi s ; y_hi = x_hi c + x_lo s   =>   dx_lo = dy_lo c + dy_hi s ; dx_hi = dy_hi c - dy_lo s
      const int fi = d < r2 ? d : d - r2;
      const float ang = (float)t * inv_freq[fi];
      float sn, cs;
      sincosf(ang, &sn, &cs);
      const float partner = d < r2 ? dqkv[i + r2] : -dqkv[i - r2];
      x = x * cs + partner * sn;
    }
    out[i] = __float2bfloat16(x);
  }
}
int rotary_bwd_to_bf16(const float* dqkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                       cudaStream_t stream) {
  if (rot > dh || (rot & 1) || rows <= 0) return NUWA_ERR_INVALID;
  const long long total = (long long)rows * 3 * H * dh;
  rotary_bwd_kernel<<<ew_grid(total), 256, 0, stream>>>(dqkv, reinterpret_cast<bf16*>(out), inv_freq, rows, n, H, dh, rot);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// dst[map[r]][0:cols] (+)= src[r][0:cols]   (map NULL = identity, negative entries are skipped)
// =================================================================================================
__global__ void __launch_bounds__(256)
add_rows_kernel(float* __restrict__ dst, long long ld_dst, const float* __restrict__ src, long long ld_src,
                const int* __restrict__ map, int rows, int cols, int accumulate) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const int dr = map ? map[r] : r;
    if (dr < 0) continue;
    float* d = dst + (long long)dr * ld_dst + c;
    const float v = src[(long long)r * ld_src + c];
    *d = accumulate ? *d + v : v;
  }
}
int add_rows_f32(float* dst, long long ld_dst, const float* src, long long ld_src, const int* map, int rows, int cols,
                 int accumulate, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return NUWA_ERR_INVALID;
  add_rows_kernel<<<ew_grid((long long)rows * cols), 256, 0, stream>>>(dst, ld_dst, src, ld_src, map, rows, cols, accumulate);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// Attention backward, shared row kernel.
//   inputs : S   fp32 [B][H][nq][jp]   logits (already scaled; masked slots = -FLT_MAX; pad slots j >= J ignored)
//            dPp fp32 [B][H][nq][jp]   d(loss)/d(P') = dO . V^T   (P' = talking-heads mix of the softmax P)
//   outputs: Pp  bf16 [B][H][nq][jp]   P'                      (operand of dV = P'^T dO)
//            dS  bf16 [B][H][nq][jp]   dS * out_scale          (operand of dQ = dS K, dK = dS^T Q)
//            dtalk fp32 [H][H]        += sum dP'[g][j] P[h][j]
//   pad columns [J, jp) of the outputs are written as zeros.
// One warp per (b, q); the H x J rows of a query live in shared memory.
// =================================================================================================
template <int H>
__global__ void __launch_bounds__(128) attn_bwd_rows_kernel(const nuwa_attn_rows_params p) {
  extern __shared__ float sm_rows[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int J = p.J, jp = p.jp;
  float* P = sm_rows + (size_t)warp * 2 * H * jp;   // [H][jp]
  float* G = P + (size_t)H * jp;                    // [H][jp]  dP' then dP
  float* Wt = sm_rows + (size_t)wpb * 2 * H * jp;   // [H][H]
  __shared__ float dW_cta[H * H];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    Wt[i] = p.talk != nullptr ? p.talk[i] : ((i / H) == (i % H) ? 1.0f : 0.0f);
    dW_cta[i] = 0.f;
  }
  __syncthreads();
  float dW[H * H];
#pragma unroll
  for (int i = 0; i < H * H; ++i) dW[i] = 0.f;
  const long long nrows = (long long)p.B * p.nq;
  const long long hs = (long long)p.nq * jp;  // head stride
  for (long long r = blockIdx.x * (long long)wpb + warp; r < nrows; r += (long long)gridDim.x * wpb) {
    const int b = (int)(r / p.nq), q = (int)(r - (long long)b * p.nq);
    const long long base = ((long long)b * H * p.nq + q) * jp;
    // ---- softmax per head ----
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float* s = p.S + base + h * hs;
      const float* d = p.dPp + base + h * hs;
      float m = -FLT_MAX;
      for (int j = lane; j < J; j += 32) {
        const float v = s[j];
        P[h * jp + j] = v;
        G[h * jp + j] = d[j];
        m = fmaxf(m, v);
      }
      m = warp_max(m);
      float sum = 0.f;
      for (int j = lane; j < J; j += 32) {
        const float e = __expf(P[h * jp + j] - m);
        P[h * jp + j] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j < J; j += 32) P[h * jp + j] *= inv;
    }
    __syncwarp();
    // ---- talking heads forward (P') and backward (dP = W^T dP'), dW accumulation ----
    for (int j = lane; j < jp; j += 32) {
      float pin[H], dpp[H];
      if (j < J) {
#pragma unroll
        for (int h = 0; h < H; ++h) { pin[h] = P[h * jp + j]; dpp[h] = G[h * jp + j]; }
      } else {
#pragma unroll
        for (int h = 0; h < H; ++h) { pin[h] = 0.f; dpp[h] = 0.f; }
      }
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          a = fmaf(Wt[g * H + h], pin[h], a);
          dW[g * H + h] = fmaf(dpp[g], pin[h], dW[g * H + h]);
        }
        reinterpret_cast<bf16*>(p.Pp)[base + g * hs + j] = __float2bfloat16(a);
      }
      if (j < J) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float a = 0.f;
#pragma unroll
          for (int g = 0; g < H; ++g) a = fmaf(Wt[g * H + h], dpp[g], a);
          G[h * jp + j] = a;  // dP[h][j]
        }
      }
    }
    __syncwarp();
    // ---- softmax backward ----
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float dot = 0.f;
      for (int j = lane; j < J; j += 32) dot += P[h * jp + j] * G[h * jp + j];
      dot = warp_sum(dot);
      bf16* o = reinterpret_cast<bf16*>(p.dS) + base + h * hs;
      for (int j = lane; j < jp; j += 32) {
        const float v = j < J ? P[h * jp + j] * (G[h * jp + j] - dot) * p.out_scale : 0.f;
        o[j] = __float2bfloat16(v);
      }
    }
    __syncwarp();
  }
  if (p.dtalk != nullptr) {
#pragma unroll
    for (int i = 0; i < H * H; ++i) {
      const float v = warp_sum(dW[i]);
      if (lane == 0) atomicAdd(&dW_cta[i], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) atomicAdd(p.dtalk + i, dW_cta[i]);
  }
}
// Register-resident variant: lane l owns the key slots j = l + 32 i (i < NJ) of ALL heads, so the softmax statistics are
// warp shuffles (the H reductions are independent -> instruction-level parallelism), the talking-heads mixes are pure
// register FMAs, and nothing but the H x H mixing matrix lives in shared memory.  J <= 32 * NJ.
template <int H, int NJ>
__global__ void __launch_bounds__(128, 2) attn_bwd_rows_reg_kernel(const nuwa_attn_rows_params p) {
  __shared__ __align__(16) float Wt[H * H];
  __shared__ float dW_cta[H * H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int J = p.J, jp = p.jp;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    Wt[i] = p.talk != nullptr ? p.talk[i] : ((i / H) == (i % H) ? 1.0f : 0.0f);
    dW_cta[i] = 0.f;
  }
  __syncthreads();
  float dW[H * H];
#pragma unroll
  for (int i = 0; i < H * H; ++i) dW[i] = 0.f;
  const long long nrows = (long long)p.B * p.nq;
  const long long hs = (long long)p.nq * jp;  // head stride
  bf16* Pp = reinterpret_cast<bf16*>(p.Pp);
  bf16* dS = reinterpret_cast<bf16*>(p.dS);
  for (long long r = blockIdx.x * (long long)wpb + warp; r < nrows; r += (long long)gridDim.x * wpb) {
    const int b = (int)(r / p.nq), q = (int)(r - (long long)b * p.nq);
    const long long base = ((long long)b * H * p.nq + q) * jp;
    float P[H][NJ], G[H][NJ];
    bool live[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = lane + 32 * i;
      live[i] = j < J;
      if (live[i] && p.key_mask != nullptr) {
        const int key = j - p.has_null;
        if (key >= 0 && p.key_mask[(long long)b * p.mask_bs + key] == 0) live[i] = false;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        P[h][i] = live[i] ? p.S[base + h * hs + j] : -FLT_MAX;
        G[h][i] = live[i] ? p.dPp[base + h * hs + j] : 0.f;
      }
    }
    // ---- softmax of every head ----
    float m[H], sum[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      m[h] = P[h][0];
#pragma unroll
      for (int i = 1; i < NJ; ++i) m[h] = fmaxf(m[h], P[h][i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < H; ++h) m[h] = fmaxf(m[h], __shfl_xor_sync(0xffffffffu, m[h], o));
#pragma unroll
    for (int h = 0; h < H; ++h) {
      sum[h] = 0.f;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        P[h][i] = live[i] ? __expf(P[h][i] - m[h]) : 0.f;
        sum[h] += P[h][i];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < H; ++h) sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], o);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float inv = 1.0f / sum[h];
#pragma unroll
      for (int i = 0; i < NJ; ++i) P[h][i] *= inv;
    }
    // ---- talking heads forward (P' -> HBM), backward (dP = W^T dP'), dW accumulation; then P .* dP row sums ----
    float dot[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dot[h] = 0.f;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = lane + 32 * i;
      float dp[H];
#pragma unroll
      for (int h = 0; h < H; ++h) dp[h] = 0.f;
#pragma unroll
      for (int g = 0; g < H; ++g) {
        float wrow[H];
#pragma unroll
        for (int h = 0; h < H; ++h) wrow[h] = Wt[g * H + h];
        const float dpp = G[g][i];
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          a = fmaf(wrow[h], P[h][i], a);
          dp[h] = fmaf(wrow[h], dpp, dp[h]);
          dW[g * H + h] = fmaf(dpp, P[h][i], dW[g * H + h]);
        }
        if (j < jp) Pp[base + g * hs + j] = __float2bfloat16(a);
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        G[h][i] = dp[h];
        dot[h] = fmaf(P[h][i], dp[h], dot[h]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < H; ++h) dot[h] += __shfl_xor_sync(0xffffffffu, dot[h], o);
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = lane + 32 * i;
      if (j < jp) {
#pragma unroll
        for (int h = 0; h < H; ++h) dS[base + h * hs + j] = __float2bfloat16(P[h][i] * (G[h][i] - dot[h]) * p.out_scale);
      }
    }
  }
  if (p.dtalk != nullptr) {
#pragma unroll
    for (int i = 0; i < H * H; ++i) {
      const float v = warp_sum(dW[i]);
      if (lane == 0) atomicAdd(&dW_cta[i], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) atomicAdd(p.dtalk + i, dW_cta[i]);
  }
}

// Two warps per row (lane l of warp half w owns the key slots j = 32 w + l + 64 i, i < NJ): for 257-slot rows the
// one-warp variant above needs P[8][9] + dP'[8][9] + 64 dW accumulators per thread and spills 768 bytes; here a thread
// holds 5 slots per head (80 + 64 registers), twice as many warps are in flight per row, and the three row statistics
// (max, sum, sum_j P dP) cross the warp pair through shared memory + a 64-thread named barrier.
template <int H, int NJ>
__global__ void __launch_bounds__(128, 2) attn_bwd_rows_pair_kernel(const nuwa_attn_rows_params p) {
  __shared__ __align__(16) float Wt[H * H];
  __shared__ float dW_cta[H * H];
  __shared__ float xch[2][3][2][H];  // [pair][statistic][warp half][head]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = warp & 1;
  const int J = p.J, jp = p.jp;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    Wt[i] = p.talk != nullptr ? p.talk[i] : ((i / H) == (i % H) ? 1.0f : 0.0f);
    dW_cta[i] = 0.f;
  }
  __syncthreads();
  const uint32_t wt_u = smem_u32(Wt);
  float dW[H * H];
#pragma unroll
  for (int i = 0; i < H * H; ++i) dW[i] = 0.f;
  const long long nrows = (long long)p.B * p.nq;
  const long long hs = (long long)p.nq * jp;  // head stride
  bf16* Pp = reinterpret_cast<bf16*>(p.Pp);
  bf16* dS = reinterpret_cast<bf16*>(p.dS);
  // combine a per-head statistic over the two warps of the pair (every lane ends up with the row value)
  auto pair_reduce = [&](float (&v)[H], int which, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float u = __shfl_xor_sync(0xffffffffu, v[h], o);
        v[h] = is_max ? fmaxf(v[h], u) : v[h] + u;
      }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) xch[pair][which][half][h] = v[h];
    }
    if (pair == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else asm volatile("bar.sync 2, 64;" ::: "memory");
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float u = xch[pair][which][half ^ 1][h];
      v[h] = is_max ? fmaxf(v[h], u) : v[h] + u;
    }
  };
  for (long long r = blockIdx.x * 2LL + pair; r < nrows; r += (long long)gridDim.x * 2) {
    const int b = (int)(r / p.nq), q = (int)(r - (long long)b * p.nq);
    const long long base = ((long long)b * H * p.nq + q) * jp;
    float P[H][NJ], G[H][NJ];
    bool live[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = half * 32 + lane + 64 * i;
      live[i] = j < J;
      if (live[i] && p.key_mask != nullptr) {
        const int key = j - p.has_null;
        if (key >= 0 && p.key_mask[(long long)b * p.mask_bs + key] == 0) live[i] = false;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        P[h][i] = live[i] ? p.S[base + h * hs + j] : -FLT_MAX;
        G[h][i] = live[i] ? p.dPp[base + h * hs + j] : 0.f;
      }
    }
    // ---- softmax of every head ----
    float m[H], sum[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      m[h] = P[h][0];
#pragma unroll
      for (int i = 1; i < NJ; ++i) m[h] = fmaxf(m[h], P[h][i]);
    }
    pair_reduce(m, 0, true);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      sum[h] = 0.f;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        P[h][i] = live[i] ? __expf(P[h][i] - m[h]) : 0.f;
        sum[h] += P[h][i];
      }
    }
    pair_reduce(sum, 1, false);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float inv = 1.0f / sum[h];
#pragma unroll
      for (int i = 0; i < NJ; ++i) P[h][i] *= inv;
    }
    // ---- talking heads forward (P' -> HBM), backward (dP = W^T dP'), dW accumulation; then P .* dP row sums ----
    float dot[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dot[h] = 0.f;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = half * 32 + lane + 64 * i;
      float dp[H];
#pragma unroll
      for (int h = 0; h < H; ++h) dp[h] = 0.f;
#pragma unroll
      for (int g = 0; g < H; ++g) {
        const float dpp = G[g][i];
        float a = 0.f;
        float wrow[H];  // re-read from shared memory per (slot, g): hoisted into registers the 64 weights spill
#pragma unroll
        for (int h4 = 0; h4 < H; h4 += 4) {
          if constexpr (H % 4 == 0) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(wrow[h4]), "=f"(wrow[h4 + 1]), "=f"(wrow[h4 + 2]), "=f"(wrow[h4 + 3])
                         : "r"(wt_u + (uint32_t)(g * H + h4) * 4u));
          } else {
            for (int h = h4; h < H && h < h4 + 4; ++h) wrow[h] = Wt[g * H + h];
          }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float w = wrow[h];
          a = fmaf(w, P[h][i], a);
          dp[h] = fmaf(w, dpp, dp[h]);
          dW[g * H + h] = fmaf(dpp, P[h][i], dW[g * H + h]);
        }
        if (j < jp) Pp[base + g * hs + j] = __float2bfloat16(a);
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        G[h][i] = dp[h];
        dot[h] = fmaf(P[h][i], dp[h], dot[h]);
      }
    }
    pair_reduce(dot, 2, false);
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const int j = half * 32 + lane + 64 * i;
      if (j < jp) {
#pragma unroll
        for (int h = 0; h < H; ++h) dS[base + h * hs + j] = __float2bfloat16(P[h][i] * (G[h][i] - dot[h]) * p.out_scale);
      }
    }
  }
  if (p.dtalk != nullptr) {
#pragma unroll
    for (int i = 0; i < H * H; ++i) {
      const float v = warp_sum(dW[i]);
      if (lane == 0) atomicAdd(&dW_cta[i], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) atomicAdd(p.dtalk + i, dW_cta[i]);
  }
}

template <int H>
static bool launch_rows_reg(const nuwa_attn_rows_params& p, int grid, cudaStream_t stream) {
  const int nj = ceil_div(p.J, 32);
  if (nj <= 1) attn_bwd_rows_reg_kernel<H, 1><<<grid, 128, 0, stream>>>(p);
  else if (nj <= 2) attn_bwd_rows_reg_kernel<H, 2><<<grid, 128, 0, stream>>>(p);
  else if (nj <= 4) attn_bwd_rows_reg_kernel<H, 4><<<grid, 128, 0, stream>>>(p);
  else if (nj <= 10 && H * 5 <= 40 && H <= 32) attn_bwd_rows_pair_kernel<H, 5><<<grid, 128, 0, stream>>>(p);  // two warps per row
  else if (nj <= 9 && H * 9 <= 72) attn_bwd_rows_reg_kernel<H, 9><<<grid, 128, 0, stream>>>(p);
  else return false;
  return true;
}

int attn_bwd_rows(const nuwa_attn_rows_params& p, cudaStream_t stream) {
  if (p.B <= 0 || p.nq <= 0 || p.J <= 0 || p.jp < p.J) return NUWA_ERR_INVALID;
  {
    const long long nrows = (long long)p.B * p.nq;
    long long want = (nrows + 4 * 8 - 1) / (4 * 8);  // >= 8 rows per warp amortise the dW reduction
    const long long cap = (long long)device_sm_count() * 2;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    bool ok = false;
    switch (p.H) {
      case 8: ok = launch_rows_reg<8>(p, grid, stream); break;
      case 4: ok = launch_rows_reg<4>(p, grid, stream); break;
      case 2: ok = launch_rows_reg<2>(p, grid, stream); break;
      case 1: ok = launch_rows_reg<1>(p, grid, stream); break;
      default: break;
    }
    if (ok) {
      NUWA_CHECK_LAUNCH();
      return NUWA_OK;
    }
  }
  if (p.key_mask != nullptr) return NUWA_ERR_INVALID;  // the shared-memory fallback expects masks folded into S
  const int wpb = 4;
  const size_t smem = ((size_t)wpb * 2 * p.H * p.jp + p.H * p.H) * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  const long long nrows = (long long)p.B * p.nq;
  long long want = (nrows + wpb * 4 - 1) / (wpb * 4);
  const long long cap = (long long)device_sm_count() * 8;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
#define NUWA_ROWS(HH)                                                                                                  \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      cudaFuncSetAttribute(attn_bwd_rows_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    attn_bwd_rows_kernel<HH><<<grid, wpb * 32, smem, stream>>>(p);                                                     \
  } while (0)
  switch (p.H) {
    case 8: NUWA_ROWS(8); break;
    case 4: NUWA_ROWS(4); break;
    case 2: NUWA_ROWS(2); break;
    case 1: NUWA_ROWS(1); break;
    default: return NUWA_ERR_INVALID;
  }
#undef NUWA_ROWS
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// =================================================================================================
// dense attention: keys/values with the learned null slot prepended, padded to jp rows
//   kfull/vfull bf16 [B][jp][inner] : row 0 = null_k / null_v, rows 1..nk = k / v, rows > nk = 0
// =================================================================================================
__global__ void __launch_bounds__(256)
kv_full_build_kernel(const bf16* __restrict__ k, const bf16* __restrict__ v, long long kv_bs, int kv_rs,
                     const float* __restrict__ null_k, const float* __restrict__ null_v, bf16* __restrict__ kfull,
                     bf16* __restrict__ vfull, int B, int nk, int jp, int inner, int has_null) {
  const long long total = (long long)B * jp * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % inner);
    const int j = (int)((i / inner) % jp);
    const int b = (int)(i / ((long long)inner * jp));
    bf16 kk = __float2bfloat16(0.f), vv = kk;
    const int src = j - has_null;
    if (has_null && j == 0) {
      kk = __float2bfloat16(null_k[c]);
      vv = __float2bfloat16(null_v[c]);
    } else if (src >= 0 && src < nk) {
      kk = k[(long long)b * kv_bs + (long long)src * kv_rs + c];
      vv = v[(long long)b * kv_bs + (long long)src * kv_rs + c];
    }
    kfull[i] = kk;
    vfull[i] = vv;
  }
}
int kv_full_build(const void* k, const void* v, long long kv_bs, int kv_rs, const float* null_k, const float* null_v,
                  void* kfull, void* vfull, int B, int nk, int jp, int inner, cudaStream_t stream) {
  const int has_null = null_k != nullptr;
  if (B <= 0 || nk <= 0 || jp < nk + has_null || inner <= 0) return NUWA_ERR_INVALID;
  kv_full_build_kernel<<<ew_grid((long long)B * jp * inner), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), kv_bs, kv_rs, null_k, null_v,
      reinterpret_cast<bf16*>(kfull), reinterpret_cast<bf16*>(vfull), B, nk, jp, inner, has_null);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
// dkfull/dvfull fp32 [B][jp][inner] -> dnull_k/dnull_v (+= over the batch, row 0) and dk/dv rows (bf16 or fp32, strided)
__global__ void __launch_bounds__(256)
kv_full_split_kernel(const float* __restrict__ dkfull, const float* __restrict__ dvfull, float* __restrict__ dnull_k,
                     float* __restrict__ dnull_v, bf16* __restrict__ dk16, bf16* __restrict__ dv16, float* __restrict__ dk32,
                     float* __restrict__ dv32, long long o_bs, int o_rs, int B, int nk, int jp, int inner, int has_null) {
  const long long total = (long long)B * (nk + has_null) * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % inner);
    const int j = (int)((i / inner) % (nk + has_null));
    const int b = (int)(i / ((long long)inner * (nk + has_null)));
    const long long src = ((long long)b * jp + j) * inner + c;
    const float gk = dkfull[src], gv = dvfull[src];
    if (has_null && j == 0) {
      atomicAdd(dnull_k + c, gk);
      atomicAdd(dnull_v + c, gv);
    } else {
      const long long o = (long long)b * o_bs + (long long)(j - has_null) * o_rs + c;
      if (dk16 != nullptr) { dk16[o] = __float2bfloat16(gk); dv16[o] = __float2bfloat16(gv); }
      else { dk32[o] = gk; dv32[o] = gv; }
    }
  }
}
int kv_full_split(const float* dkfull, const float* dvfull, float* dnull_k, float* dnull_v, void* dk16, void* dv16,
                  float* dk32, float* dv32, long long o_bs, int o_rs, int B, int nk, int jp, int inner,
                  cudaStream_t stream) {
  const int has_null = dnull_k != nullptr;
  if (B <= 0 || nk <= 0 || inner <= 0) return NUWA_ERR_INVALID;
  if ((dk16 == nullptr) == (dk32 == nullptr)) return NUWA_ERR_INVALID;
  kv_full_split_kernel<<<ew_grid((long long)B * (nk + has_null) * inner), 256, 0, stream>>>(
      dkfull, dvfull, dnull_k, dnull_v, reinterpret_cast<bf16*>(dk16), reinterpret_cast<bf16*>(dv16), dk32, dv32, o_bs,
      o_rs, B, nk, jp, inner, has_null);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// dense key mask applied to the logits: S[b][h][q][j] = -FLT_MAX where key j is masked (slot 0 = null key is never masked)
__global__ void __launch_bounds__(256)
mask_scores_kernel(float* __restrict__ S, const unsigned char* __restrict__ mask, int mask_bs, int B, int H, int nq, int jp,
                   int nk, int has_null) {
  const long long total = (long long)B * H * nq * jp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % jp);
    const int b = (int)(i / ((long long)H * nq * jp));
    const int key = j - has_null;
    if (key >= 0 && key < nk && mask[(long long)b * mask_bs + key] == 0) S[i] = -FLT_MAX;
  }
}
int mask_scores(float* S, const unsigned char* mask, int mask_bs, int B, int H, int nq, int jp, int nk, int has_null,
                cudaStream_t stream) {
  if (mask == nullptr) return NUWA_OK;
  mask_scores_kernel<<<ew_grid((long long)B * H * nq * jp), 256, 0, stream>>>(S, mask, mask_bs, B, H, nq, jp, nk, has_null);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
