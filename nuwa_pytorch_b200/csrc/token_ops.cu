// Token-level kernels around the transformer stacks: embedding + axial positions + bos
// (nuwa_pytorch.py:1693-1709,1940-1944), rotary embedding of q,k,v (nuwa_pytorch.py:132-153,333-335),
// cross entropy over the logits (nuwa_pytorch.py:1963) and the generate() sampling step: classifier-free
// guidance mix, top-k filter, Gumbel arg-max (nuwa_pytorch.py:55-66,1713-1719,1901-1906).
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

// out[b, tl, :] for absolute position t = t0 + tl:
//   has_bos && t == 0 : bos vector
//   else              : table[idx[b, t - has_bos]] + ax1[p / (d2*d3)] + ax2[(p / d3) % d2] + ax3[p % d3],  p = t - has_bos
__global__ void __launch_bounds__(256) embed_kernel(const EmbedParams p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.nt) return;
  const int b = row / p.nt, tl = row - b * p.nt;
  const int t = (p.t0_ptr != nullptr ? __ldg(p.t0_ptr) : p.t0) + tl;
  float* out = p.out + (long long)row * p.D;
  if (p.has_bos && t == 0) {
    for (int c = lane * 4; c < p.D; c += 128)
      *reinterpret_cast<float4*>(out + c) = *reinterpret_cast<const float4*>(p.bos + c);
    return;
  }
  const int pidx = t - (p.has_bos ? 1 : 0);
  const long long tok = p.idx[(long long)b * p.idx_bs + pidx];
  const float* e = p.table + tok * p.D;
  const float* a1 = p.ax1 ? p.ax1 + (long long)(pidx / (p.d2 * p.d3)) * p.D : nullptr;
  const float* a2 = p.ax2 ? p.ax2 + (long long)((pidx / p.d3) % p.d2) * p.D : nullptr;
  const float* a3 = p.ax3 ? p.ax3 + (long long)(pidx % p.d3) * p.D : nullptr;
  for (int c = lane * 4; c < p.D; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(e + c);
    // reference order: ((a1 + a2) + a3) + embedding  (nuwa_pytorch.py:1704,1941)
    float4 pos = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
    if (a1) { pos = *reinterpret_cast<const float4*>(a1 + c); any = true; }
    if (a2) {
      const float4 u = *reinterpret_cast<const float4*>(a2 + c);
      if (any) { pos.x += u.x; pos.y += u.y; pos.z += u.z; pos.w += u.w; } else { pos = u; any = true; }
    }
    if (a3) {
      const float4 u = *reinterpret_cast<const float4*>(a3 + c);
      if (any) { pos.x += u.x; pos.y += u.y; pos.z += u.z; pos.w += u.w; } else { pos = u; any = true; }
    }
    if (any) { v.x = pos.x + v.x; v.y = pos.y + v.y; v.z = pos.z + v.z; v.w = pos.w + v.w; }
    *reinterpret_cast<float4*>(out + c) = v;
  }
}

int embed_tokens(const EmbedParams& p, cudaStream_t stream) {
  if (p.D % 4 != 0 || p.B <= 0 || p.nt <= 0) return NUWA_ERR_INVALID;
  const int wpb = 8;
  embed_kernel<<<ceil_div(p.B * p.nt, wpb), wpb * 32, 0, stream>>>(p);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// qkv fp32 [B*n, 3*inner] -> bf16, rotating the first rot dims of every head of q, k AND v (D10).
__global__ void __launch_bounds__(256)
rotary_kernel(const float* __restrict__ qkv, bf16* __restrict__ out, const float* __restrict__ inv_freq, int rows,
              int n, int H, int dh, int rot) {
  const long long total = (long long)rows * 3 * H * dh;
  const int r2 = rot / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dh);
    const long long row = i / (3LL * H * dh);
    const int t = (int)(row % n);
    float x = qkv[i];
    if (d < rot) {
      const int fi = d < r2 ? d : d - r2;
      const float ang = (float)t * inv_freq[fi];
      float sn, cs;
      sincosf(ang, &sn, &cs);
      const float partner = d < r2 ? -qkv[i + r2] : qkv[i - r2];
      x = x * cs + partner * sn;
    }
    out[i] = __float2bfloat16(x);
  }
}

int rotary_to_bf16(const float* qkv, void* out, const float* inv_freq, int rows, int n, int H, int dh, int rot,
                   cudaStream_t stream) {
  if (rot > dh || (rot & 1) || rows <= 0) return NUWA_ERR_INVALID;
  const long long total = (long long)rows * 3 * H * dh;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  rotary_kernel<<<grid, 256, 0, stream>>>(qkv, reinterpret_cast<bf16*>(out), inv_freq, rows, n, H, dh, rot);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// cross entropy: per-row  logsumexp(logits) - logits[target] ; then a deterministic mean
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ row_loss,
               int rows, int V, int ld) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const float* l = logits + (long long)row * ld;
  __shared__ float red[32];
  float m = -FLT_MAX;
  for (int c = threadIdx.x; c < V; c += blockDim.x) m = fmaxf(m, l[c]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) s += expf(l[c] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
    row_loss[row] = (m + logf(tot)) - l[target[row]];
  }
}
// Same, the row held in registers (V <= 256 * 4 * NV, 16-byte aligned rows): ONE pass over HBM with 16-byte loads
// (the two-pass kernel above reaches 1.7 TB/s on the 20480 x 8192 logits of cfg 3: scalar loads, 1 row per CTA pass).
template <int NV>
__global__ void __launch_bounds__(256)
ce_rows_reg_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ row_loss,
                   int rows, int V, int ld) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const float4* l4 = reinterpret_cast<const float4*>(logits + (long long)row * ld);
  const int nv4 = V >> 2;
  __shared__ float red[2][8];
  float4 v[NV];
  float m = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = c < nv4 ? __ldcs(l4 + c) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
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
  for (int i = 0; i < NV; ++i) s += (__expf(v[i].x - m) + __expf(v[i].y - m)) + (__expf(v[i].z - m) + __expf(v[i].w - m));
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[1][i];
    row_loss[row] = (m + logf(tot)) - logits[(long long)row * ld + target[row]];
  }
}
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    out[0] = (float)(t / (double)n);
  }
}

int cross_entropy_mean(const float* logits, int ld, const long long* target, float* row_loss, float* out, int rows,
                       int V, cudaStream_t stream) {
  if (rows <= 0 || V <= 0) return NUWA_ERR_INVALID;
  const bool vec = (V % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec && V <= 256 * 4 * 2) ce_rows_reg_kernel<2><<<rows, 256, 0, stream>>>(logits, target, row_loss, rows, V, ld);
  else if (vec && V <= 256 * 4 * 8) ce_rows_reg_kernel<8><<<rows, 256, 0, stream>>>(logits, target, row_loss, rows, V, ld);
  else ce_rows_kernel<<<rows, 256, 0, stream>>>(logits, target, row_loss, rows, V, ld);
  NUWA_CHECK_LAUNCH();
  mean_kernel<<<1, 1024, 0, stream>>>(row_loss, out, rows);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// sampling step of generate(): one CTA per batch row.
//   l = u + (c - u) * cond_scale ; keep the k largest (others -inf) ; argmax(l / temperature + gumbel(noise))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {  // order-preserving float -> uint
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, const float* __restrict__ noise,
              long long* __restrict__ out, float* __restrict__ guided_out, int V, int k, float cond_scale,
              float temperature, const int* __restrict__ step_ptr, long long out_bs) {
  extern __shared__ float sl[];  // V guided logits
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  if (step_ptr != nullptr) {  // graph-replayed decode step: this step's noise slab and output column
    const int step = __ldg(step_ptr);
    noise += (long long)step * gridDim.x * V;
    out += (long long)b * out_bs + step - b;  // so that out[b] below lands on out0[b*out_bs + step]
  }
  const float* c = cond + (long long)b * V;
  const float* u = uncond ? uncond + (long long)b * V : nullptr;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float l = c[i];
    if (u) l = u[i] + (l - u[i]) * cond_scale;
    sl[i] = l;
    if (guided_out) guided_out[(long long)b * V + i] = l;
  }
  if (threadIdx.x == 0) { s_prefix = 0; s_remaining = (uint32_t)k; }
  __syncthreads();
  // radix select (MSB first) of the k-th largest ordered key
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const uint32_t key = f2ord(sl[i]);
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t rem = s_remaining;
      int d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      s_prefix = prefix | ((uint32_t)d << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;  // ordered key of the k-th largest logit
  float best = -FLT_MAX;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float l = sl[i];
    if (f2ord(l) < kth) continue;  // filtered to -inf
    const float un = noise[(long long)b * V + i];
    const float g = -logf(fmaxf(-logf(fmaxf(un, 1e-20f)), 1e-20f));
    const float s = l / temperature + g;
    if (s > best || (s == best && i < besti)) { best = s; besti = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      if (s_val[w] > best || (s_val[w] == best && s_idx[w] < besti)) { best = s_val[w]; besti = s_idx[w]; }
    out[b] = besti;
  }
}

int sample_topk_gumbel(const float* cond, const float* uncond, const float* noise, long long* out, float* guided_out,
                       int B, int V, int k, float cond_scale, float temperature, cudaStream_t stream) {
  if (B <= 0 || V <= 0 || k <= 0 || k > V) return NUWA_ERR_INVALID;
  const size_t smem = (size_t)V * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  if (smem > 48 * 1024) cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_kernel<<<B, 256, smem, stream>>>(cond, uncond, noise, out, guided_out, V, k, cond_scale, temperature, nullptr, 0);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

int sample_topk_gumbel_at(const float* cond, const float* uncond, const float* noise, long long* out, long long out_bs,
                          const int* step_ptr, int B, int V, int k, float cond_scale, float temperature,
                          cudaStream_t stream) {
  if (B <= 0 || V <= 0 || k <= 0 || k > V || step_ptr == nullptr) return NUWA_ERR_INVALID;
  const size_t smem = (size_t)V * sizeof(float);
  if (smem > 200 * 1024) return NUWA_ERR_INVALID;
  if (smem > 48 * 1024) cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_kernel<<<B, 256, smem, stream>>>(cond, uncond, noise, out, nullptr, V, k, cond_scale, temperature, step_ptr, out_bs);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// cache[b][*t][0:width] = row[b][0:width]   (bf16; appends the decoded token's q|k|v to the KV cache)
__global__ void __launch_bounds__(256)
cache_append_kernel(const bf16* __restrict__ row, bf16* __restrict__ cache, long long cache_bs, int width, int B,
                    const int* __restrict__ t_ptr) {
  const int t = __ldg(t_ptr);
  const int w8 = width / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * w8; i += gridDim.x * blockDim.x) {
    const int b = i / w8, c = (i - b * w8) * 8;
    *reinterpret_cast<uint4*>(cache + (long long)b * cache_bs + (long long)t * width + c) =
        *reinterpret_cast<const uint4*>(row + (long long)b * width + c);
  }
}
int cache_append(const void* row, void* cache, long long cache_bs, int width, int B, const int* t_ptr, cudaStream_t stream) {
  if (B <= 0 || width <= 0 || (width % 8) || t_ptr == nullptr) return NUWA_ERR_INVALID;
  const int n = B * (width / 8);
  cache_append_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(row), reinterpret_cast<bf16*>(cache),
                                                            cache_bs, width, B, t_ptr);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
__global__ void step_increment_kernel(int* t) { *t += 1; }
int step_increment(int* t_ptr, cudaStream_t stream) {
  if (t_ptr == nullptr) return NUWA_ERR_INVALID;
  step_increment_kernel<<<1, 1, 0, stream>>>(t_ptr);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
