// VQGanVAE kernels around the tensor-core convolutions (vqgan_vae.py): layout changes at the API
// boundary (NCHW fp32 <-> NHWC bf16), im2col of the few-channel first convolution, GroupNorm(+LeakyReLU),
// bilinear 2x upsampling, the spatial-axis l2norm of VQGanAttention, the VQ codebook arg-max with a
// warp-shuffle reduction, codebook gather and the final dim->channels 1x1 convolution.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

// ------------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 (tiled transpose through shared memory)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, bf16* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? in[((long long)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) out[((long long)b * HW + p) * C + c] = __float2bfloat16(tile[tx][i]);
  }
}
int nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return NUWA_ERR_INVALID;
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, stream>>>(in, reinterpret_cast<bf16*>(out), C, H * W);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// NHWC (fp32 or bf16) -> NCHW fp32
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      if constexpr (sizeof(T) == 2) v = __bfloat162float(in[((long long)b * HW + p) * C + c]);
      else v = in[((long long)b * HW + p) * C + c];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) out[((long long)b * C + c) * HW + p] = tile[tx][i];
  }
}
int nhwc_to_nchw_f32(const void* in, int in_is_bf16, float* out, int B, int C, int H, int W, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return NUWA_ERR_INVALID;
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B);
  if (in_is_bf16) nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), out, C, H * W);
  else nhwc_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(in), out, C, H * W);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// im2col of the first convolution (vqgan_vae.py:365): NCHW fp32 image -> [B*H*W, Kpad] bf16,
// K index = (kh*KS + kw)*C + c, zero padded ('same' padding KS/2, stride 1).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int C, int H, int W, int KS, int Kpad) {
  // one thread = 8 consecutive K entries of one output pixel -> one 16-byte store.  The (dy, dx, c) of every K entry
  // comes from a per-CTA table (no integer division in the hot loop).
  extern __shared__ int lut[];  // [Kpad]: ((dy + 64) << 20) | ((dx + 64) << 10) | c ,  -1 for the zero padding
  const int pad = KS / 2;
  const int K = KS * KS * C;
  for (int kk = threadIdx.x; kk < Kpad; kk += blockDim.x) {
    int v = -1;
    if (kk < K) {
      const int c = kk % C, tap = kk / C;
      const int kh = tap / KS, kw = tap - kh * KS;
      v = ((kh - pad + 64) << 20) | ((kw - pad + 64) << 10) | c;
    }
    lut[kk] = v;
  }
  __syncthreads();
  const int G = Kpad / 8;
  const long long total = (long long)B * H * W * G;
  const long long HW = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gk = (int)(i % G);
    const long long pix = i / G;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const long long b = pix / HW;
    const float* base = img + b * C * HW;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int l = lut[gk * 8 + e];
      float val = 0.f;
      if (l >= 0) {
        const int yy = y + ((l >> 20) & 1023) - 64, xx = x + ((l >> 10) & 1023) - 64, c = l & 1023;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = __ldg(base + c * HW + (long long)yy * W + xx);
      }
      v[e] = val;
    }
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + pix * Kpad + gk * 8) = u;
  }
}
int im2col_nchw_f32(const float* img, void* out, int B, int C, int H, int W, int KS, int Kpad, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || KS <= 0 || Kpad < KS * KS * C || (Kpad % 8)) return NUWA_ERR_INVALID;
  const long long total = (long long)B * H * W * (Kpad / 8);
  long long g = (total + 255) / 256;
  int grid = (int)(g > 148LL * 64 ? 148LL * 64 : g);
  if (C > 1023 || KS > 63) return NUWA_ERR_INVALID;
  im2col_kernel<<<grid, 256, Kpad * sizeof(int), stream>>>(img, reinterpret_cast<bf16*>(out), B, C, H, W, KS, Kpad);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm over NHWC fp32 (vqgan_vae.py:218,221,233,236): stats per (sample, group), then apply
// (+ optional LeakyReLU 0.1) writing bf16 (next conv operand) and/or fp32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gn_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, int HW, int C, int G, float eps) {
  // one CTA per (sample, group); the group's channels are a contiguous cg-wide slice of every NHWC pixel row.
  const int b = blockIdx.y, g = blockIdx.x;
  const int cg = C / G;
  const float* base = x + (long long)b * HW * C + g * cg;
  const double n = (double)HW * cg;
  __shared__ double red[16];
  __shared__ float s_mean;
  if (cg % 4 != 0) {
    // narrow groups (tiny test models): scalar path, one thread does nothing clever
    double s1 = 0.0;
    for (long long i = threadIdx.x; i < (long long)HW * cg; i += blockDim.x) s1 += (double)base[(i / cg) * C + (i % cg)];
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
      s_mean = (float)(t / n);
    }
    __syncthreads();
    const float mu = s_mean;
    double q1 = 0.0;
    for (long long i = threadIdx.x; i < (long long)HW * cg; i += blockDim.x) {
      const float d = base[(i / cg) * C + (i % cg)] - mu;
      q1 += (double)d * d;
    }
    for (int o = 16; o > 0; o >>= 1) q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
      stats[((long long)b * G + g) * 2 + 0] = mu;
      stats[((long long)b * G + g) * 2 + 1] = rsqrtf((float)(t / n) + eps);
    }
    return;
  }
  const int cg4 = cg / 4;
  const long long n4 = (long long)HW * cg4;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const long long p = i / cg4;
    const int c4 = (int)(i - p * cg4);
    const float4 v = *reinterpret_cast<const float4*>(base + p * C + c4 * 4);
    s += (v.x + v.y) + (v.z + v.w);
  }
  double sd = (double)s;
  for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    s_mean = (float)(t / n);
  }
  __syncthreads();
  const float mean = s_mean;
  float q = 0.f;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const long long p = i / cg4;
    const int c4 = (int)(i - p * cg4);
    const float4 v = *reinterpret_cast<const float4*>(base + p * C + c4 * 4);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  double qd = (double)q;
  for (int o = 16; o > 0; o >>= 1) qd += __shfl_xor_sync(0xffffffffu, qd, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = qd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    stats[((long long)b * G + g) * 2 + 0] = mean;
    stats[((long long)b * G + g) * 2 + 1] = rsqrtf((float)(t / n) + eps);
  }
}
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ w,
                const float* __restrict__ bias, bf16* __restrict__ out_bf16, float* __restrict__ out_f32, long long total,
                int HW, int C, int G, int leaky) {
  const int cg = C / G;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < total;
       i += (long long)gridDim.x * blockDim.x * 4) {
    const int c = (int)(i % C);
    const long long b = i / ((long long)HW * C);
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cg;
      const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
      float r = (o[j] - mean) * rstd * w[c + j] + bias[c + j];
      if (leaky) r = leaky01(r);
      o[j] = r;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + i) = make_float4(o[0], o[1], o[2], o[3]);
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o[0], o[1]);
      pk.y = pack_bf16x2(o[2], o[3]);
      *reinterpret_cast<uint2*>(out_bf16 + i) = pk;
    }
  }
}
int groupnorm_nhwc(const float* x, const float* w, const float* bias, float* stats_ws, void* out_bf16, float* out_f32,
                   int B, int HW, int C, int G, int leaky, cudaStream_t stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || (C % G) || (C % 4)) return NUWA_ERR_INVALID;
  dim3 g1(G, B);
  gn_stats_kernel<<<g1, 512, 0, stream>>>(x, stats_ws, HW, C, G, 1e-5f);
  NUWA_CHECK_LAUNCH();
  const long long total = (long long)B * HW * C;
  long long g = (total / 4 + 255) / 256;
  int grid = (int)(g > 148LL * 16 ? 148LL * 16 : g);
  gn_apply_kernel<<<grid, 256, 0, stream>>>(x, stats_ws, w, bias, reinterpret_cast<bf16*>(out_bf16), out_f32, total, HW,
                                             C, G, leaky);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// bilinear 2x upsample, align_corners=False (vqgan_vae.py:353), NHWC bf16 -> NHWC bf16
// src = (dst + 0.5)/2 - 0.5 clamped at 0 ; taps floor(src), floor(src)+1 (clamped) -- ATen's rule.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lerp4_bf16x8(const uint4& a, const uint4& b, const uint4& c, const uint4& d, float wa,
                                              float wb, float wc, float wd) {
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  const uint32_t* pc = reinterpret_cast<const uint32_t*>(&c);
  const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d);
  uint32_t r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = unpack_bf16x2(pa[j]), fb = unpack_bf16x2(pb[j]), fc = unpack_bf16x2(pc[j]), fd = unpack_bf16x2(pd[j]);
    r[j] = pack_bf16x2(wa * fa.x + wb * fb.x + wc * fc.x + wd * fd.x, wa * fa.y + wb * fb.y + wc * fc.y + wd * fd.y);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C) {
  // align_corners=False, scale 2: output 2i+1 = .75 in[i] + .25 in[i+1], output 2i+2 = .25 in[i] + .75 in[i+1]
  // (indices clamped at the border).  One thread = the 2x2 output block fed by inputs (i,i+1)x(j,j+1), 8 channels:
  // 4 loads + 4 stores of 16 bytes.  i in [-1, H-1], j in [-1, W-1].
  const int OW = 2 * W, OH = 2 * H;
  const int C8 = C / 8;
  const long long total = (long long)B * (H + 1) * (W + 1) * C8;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % C8);
    long long r = t / C8;
    const int j = (int)(r % (W + 1)) - 1;
    r /= (W + 1);
    const int i = (int)(r % (H + 1)) - 1;
    const long long b = r / (H + 1);
    const int i0 = i < 0 ? 0 : i, i1 = (i + 1 > H - 1) ? H - 1 : i + 1;
    const int j0 = j < 0 ? 0 : j, j1 = (j + 1 > W - 1) ? W - 1 : j + 1;
    const bf16* base = in + b * (long long)H * W * C + c8 * 8;
    const uint4 v00 = *reinterpret_cast<const uint4*>(base + ((long long)i0 * W + j0) * C);
    const uint4 v01 = *reinterpret_cast<const uint4*>(base + ((long long)i0 * W + j1) * C);
    const uint4 v10 = *reinterpret_cast<const uint4*>(base + ((long long)i1 * W + j0) * C);
    const uint4 v11 = *reinterpret_cast<const uint4*>(base + ((long long)i1 * W + j1) * C);
    bf16* ob = out + b * (long long)OH * OW * C + c8 * 8;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int oy = 2 * i + 1 + dy;
      if (oy < 0 || oy >= OH) continue;
      const float wy0 = dy == 0 ? 0.75f : 0.25f, wy1 = 1.f - wy0;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ox = 2 * j + 1 + dx;
        if (ox < 0 || ox >= OW) continue;
        const float wx0 = dx == 0 ? 0.75f : 0.25f, wx1 = 1.f - wx0;
        *reinterpret_cast<uint4*>(ob + ((long long)oy * OW + ox) * C) =
            lerp4_bf16x8(v00, v01, v10, v11, wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1);
      }
    }
  }
}
int upsample2x_nhwc_bf16(const void* in, void* out, int B, int H, int W, int C, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return NUWA_ERR_INVALID;
  const long long total = (long long)B * (H + 1) * (W + 1) * (C / 8);
  long long g = (total + 255) / 256;
  int grid = (int)(g > 148LL * 64 ? 148LL * 64 : g);
  upsample2x_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), B, H, W, C);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// VQGanAttention operand prep (vqgan_vae.py:269-273): qkv fp32 [B, n, 3*inner] -> bf16 with q and k
// l2-normalised over the SPATIAL axis n (per (b, channel); eps 1e-12 as F.normalize), v copied.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spatial_l2norm_kernel(const float* __restrict__ qkv, bf16* __restrict__ out, int n, int inner) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);  // channel in [0, 3*inner)
  const int ty = threadIdx.x >> 5;                      // 8 row groups
  const int C3 = 3 * inner;
  __shared__ float part[8][33];
  const float* base = qkv + (long long)b * n * C3;
  float s = 0.f;
  if (c < 2 * inner)
    for (int i = ty; i < n; i += 8) {
      const float v = base[(long long)i * C3 + c];
      s += v * v;
    }
  part[ty][threadIdx.x & 31] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x & 31];
  const float inv = (c < 2 * inner) ? 1.0f / fmaxf(sqrtf(tot), 1e-12f) : 1.0f;
  if (c < C3)
    for (int i = ty; i < n; i += 8)
      out[((long long)b * n + i) * C3 + c] = __float2bfloat16(base[(long long)i * C3 + c] * inv);
}
int vae_attn_prep(const float* qkv, void* out, int B, int n, int inner, cudaStream_t stream) {
  if (B <= 0 || n <= 0 || inner <= 0) return NUWA_ERR_INVALID;
  dim3 grid(ceil_div(3 * inner, 32), B);
  spatial_l2norm_kernel<<<grid, 256, 0, stream>>>(qkv, reinterpret_cast<bf16*>(out), n, inner);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// VQ codebook arg-max (third-party VectorQuantize, call site vqgan_vae.py:435).  fp32 throughout so that
// token ids are bit-exact against the fp32 reference on the same inputs.
//   cosine : score = <x/|x| , e^>            (e^ = pre-normalised codebook)
//   euclid : score = -(|x|^2 - 2<x,e> + |e|^2)
// CTA = 64 tokens x (all codes in steps of 64); 256 threads, 4x4 register tile each; running best per
// thread, then a warp-shuffle arg-max across the 16 threads that share a token; first maximum wins.
// ------------------------------------------------------------------------------------------------
static constexpr int VQ_TM = 64, VQ_TN = 128, VQ_TK = 16;

__global__ void __launch_bounds__(256)
vq_argmax_kernel(const float* __restrict__ x, const float* __restrict__ code, const float* __restrict__ code_sq,
                 long long* __restrict__ out, int M, int Kc, int D, int cosine) {
  // CTA = 64 tokens x (all codes, 128 per step); 256 threads = 16 (token groups of 4) x 16 (code groups of 8):
  // 32 FMA per 3 shared-memory 128-bit loads.
  __shared__ __align__(16) float xs[VQ_TK][VQ_TM + 4];
  __shared__ __align__(16) float cs[VQ_TK][VQ_TN + 4];
  __shared__ float xnorm[VQ_TM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: code sub-tile, ty: token sub-tile
  const int m0 = blockIdx.x * VQ_TM;
  if (tid < VQ_TM) {  // per-token scale: 1/max(|x|,eps) for cosine, |x|^2 for euclid
    const int m = m0 + tid;
    float s = 0.f;
    if (m < M)
      for (int d = 0; d < D; ++d) { const float v = x[(long long)m * D + d]; s += v * v; }
    xnorm[tid] = cosine ? 1.0f / fmaxf(sqrtf(s), 1e-12f) : s;
  }
  __syncthreads();
  float best[4];
  int besti[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = -FLT_MAX; besti[i] = 0x7fffffff; }
  for (int n0 = 0; n0 < Kc; n0 += VQ_TN) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += VQ_TK) {
      // x tile: 64 x 16, code tile: 128 x 16 (coalesced along D)
      for (int i = tid; i < VQ_TM * VQ_TK; i += 256) {
        const int r = i / VQ_TK, k = i % VQ_TK;
        const int m = m0 + r;
        float v = (m < M && k0 + k < D) ? x[(long long)m * D + k0 + k] : 0.f;
        if (cosine) v *= xnorm[r];
        xs[k][r] = v;
      }
      for (int i = tid; i < VQ_TN * VQ_TK; i += 256) {
        const int r = i / VQ_TK, k = i % VQ_TK;
        const int cn = n0 + r;
        cs[k][r] = (cn < Kc && k0 + k < D) ? code[(long long)cn * D + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VQ_TK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&cs[k][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&cs[k][tx * 8 + 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cn = n0 + tx * 8 + j;
        if (cn >= Kc) continue;
        float s = acc[i][j];
        if (!cosine) s = -(xnorm[ty * 4 + i] - 2.0f * s + code_sq[cn]);
        if (s > best[i] || (s == best[i] && cn < besti[i])) { best[i] = s; besti[i] = cn; }
      }
  }
  // arg-max across the 16 lanes (tx) that hold the same tokens: a warp = 2 token groups x 16 code groups
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best[i], o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
      if (ob > best[i] || (ob == best[i] && oi < besti[i])) { best[i] = ob; besti[i] = oi; }
    }
    const int m = m0 + ty * 4 + i;
    if (tx == 0 && m < M) out[m] = besti[i];
  }
}
int vq_argmax(const float* x, const float* code, const float* code_sq, long long* out, int M, int Kc, int D, int cosine,
              cudaStream_t stream) {
  if (M <= 0 || Kc <= 0 || D <= 0) return NUWA_ERR_INVALID;
  if (!cosine && code_sq == nullptr) return NUWA_ERR_INVALID;
  vq_argmax_kernel<<<ceil_div(M, VQ_TM), 256, 0, stream>>>(x, code, code_sq, out, M, Kc, D, cosine);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// VQ arg-max on the tensor cores, exact: (1) bf16 similarity GEMM (tcgen05) of all tokens against all codes,
// (2) per token, every code whose bf16 score lies within the PROVABLE bf16 error band of the row maximum is
// re-scored in fp32 exactly as vq_argmax_kernel does; the best exact score wins, first index on ties.
//
// Error band: operands rounded to bf16 (unit roundoff u = 2^-8), fp32 accumulation:
//   |s~_j - s_j| <= (2u + u^2) sum_d |x_d e_jd| + (fp32 accumulation) <= E := 2^-7 (1 + 2^-6) |x| max_j|e_j|
// so with t_j = s_j (cosine) or 2 s_j - |e_j|^2 (euclid: arg-max of -(|x|^2 - 2 s + |e|^2)), c = 1 resp. 2:
//   t~_j >= max_j t~_j - 2 c E  holds for the true arg-max.  For unit vectors in 512 dimensions that is ~4-16 codes
// of 8192 per token.  The fp32 CUDA-core kernel needs M*Kc*D FMAs (cfg 5: 14.5 ms of a 64 ms step, cfg 2: 5 ms of
// 126 ms); this path runs the same contraction at tensor-core rate and reads the scores twice.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_prep_kernel(const float* __restrict__ x, bf16* __restrict__ xb, int M, int D, int cosine) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const float* xr = x + (long long)m * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = xr[d]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  const float inv = cosine ? 1.0f / fmaxf(sqrtf(s), 1e-12f) : 1.0f;
  for (int d = lane; d < D; d += 32) xb[(long long)m * D + d] = __float2bfloat16(xr[d] * inv);
}

template <int NV>  // D <= 128 * NV
__global__ void __launch_bounds__(256)
vq_refine_kernel(const float* __restrict__ x, const float* __restrict__ code, const float* __restrict__ code_sq,
                 const float* __restrict__ scores, long long* __restrict__ out, int M, int Kc, int D, int cosine,
                 const float* __restrict__ emax_ptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  // this lane's slice of the token row: dims lane*4 + 128*i .. +3 (scaled as vq_argmax_kernel scales it)
  float4 xv[NV];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = lane * 4 + 128 * i;
    xv[i] = d < D ? *reinterpret_cast<const float4*>(x + (long long)m * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    sq += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
  }
  sq = warp_sum(sq);
  float xnorm = sqrtf(sq);
  if (cosine) {
    const float inv = 1.0f / fmaxf(xnorm, 1e-12f);
#pragma unroll
    for (int i = 0; i < NV; ++i) { xv[i].x *= inv; xv[i].y *= inv; xv[i].z *= inv; xv[i].w *= inv; }
    xnorm = 1.0f;
  }
  const float cfac = cosine ? 1.0f : 2.0f;
  const float emax = __ldg(emax_ptr);
  const float band = 2.0f * cfac * 0.0079345703125f * xnorm * emax + 1e-30f;  // 2 c E,  E = 2^-7 (1 + 2^-6) |x| max|e|
  const float* srow = scores + (long long)m * Kc;
  // pass 1: row maximum of the bf16-operand scores
  float tmax = -FLT_MAX;
  for (int j = lane; j < Kc; j += 32) {
    const float t = cosine ? srow[j] : 2.0f * srow[j] - code_sq[j];
    tmax = fmaxf(tmax, t);
  }
  tmax = warp_max(tmax);
  const float thr = tmax - band;
  // pass 2: exact fp32 re-score of the codes inside the band, in index order
  float best = -FLT_MAX;
  int besti = 0x7fffffff;
  for (int j0 = 0; j0 < Kc; j0 += 32) {
    const int j = j0 + lane;
    bool cand = false;
    if (j < Kc) {
      const float t = cosine ? srow[j] : 2.0f * srow[j] - code_sq[j];
      cand = t >= thr;
    }
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      const int jj = j0 + (__ffs(mask) - 1);
      mask &= mask - 1;
      const float* cr = code + (long long)jj * D;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int d = lane * 4 + 128 * i;
        if (d < D) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cr + d));
          dot = fmaf(xv[i].x, c4.x, dot); dot = fmaf(xv[i].y, c4.y, dot);
          dot = fmaf(xv[i].z, c4.z, dot); dot = fmaf(xv[i].w, c4.w, dot);
        }
      }
      dot = warp_sum(dot);
      const float sc = cosine ? dot : -(sq - 2.0f * dot + code_sq[jj]);
      if (sc > best) { best = sc; besti = jj; }  // candidates arrive in increasing index order: first maximum wins
    }
  }
  // a row holding NaN / Inf passes no threshold test: fall back to index 0 like the fp32 kernel (never out of range)
  if (lane == 0) out[m] = besti < Kc ? besti : 0;
}

// workspace: M*D bf16 (token rows, normalised for cosine) followed by M*Kc fp32 (scores), 16-byte aligned.
// Envelope: D % 8 == 0, D <= 1024, Kc % 4 == 0; NUWA_ERR_INVALID outside it (caller uses vq_argmax).
size_t vq_argmax_tc_workspace(int M, int Kc, int D) {
  const size_t a = ((size_t)M * D * 2 + 255) & ~(size_t)255;
  return a + (size_t)M * Kc * 4;
}
int vq_argmax_tc(const float* x, const float* code, const float* code_sq, const void* code_bf16, const float* emax,
                 long long* out, int M, int Kc, int D, int cosine, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (M <= 0 || Kc <= 0 || D <= 0 || code_bf16 == nullptr || workspace == nullptr || emax == nullptr) return NUWA_ERR_INVALID;
  if ((D % 8) || D > 1024 || (Kc % 4)) return NUWA_ERR_INVALID;
  if (!cosine && code_sq == nullptr) return NUWA_ERR_INVALID;
  if (ws_bytes < vq_argmax_tc_workspace(M, Kc, D)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(code) & 15) || (reinterpret_cast<uintptr_t>(code_bf16) & 15))
    return NUWA_ERR_INVALID;
  bf16* xb = reinterpret_cast<bf16*>(workspace);
  float* scores = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)M * D * 2 + 255) & ~(size_t)255));
  vq_prep_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(x, xb, M, D, cosine);
  NUWA_CHECK_LAUNCH();
  const int rc = gemm_bf16(xb, D, code_bf16, D, M, Kc, D, nullptr, nullptr, 0, scores, nullptr, Kc, ACT_NONE, 0, stream);
  if (rc != NUWA_OK) return rc;
  const int grid = ceil_div(M, 8);
  if (D <= 256) vq_refine_kernel<2><<<grid, 256, 0, stream>>>(x, code, code_sq, scores, out, M, Kc, D, cosine, emax);
  else if (D <= 512) vq_refine_kernel<4><<<grid, 256, 0, stream>>>(x, code, code_sq, scores, out, M, Kc, D, cosine, emax);
  else vq_refine_kernel<8><<<grid, 256, 0, stream>>>(x, code, code_sq, scores, out, M, Kc, D, cosine, emax);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32-faithful linear on the tensor cores: VectorQuantize.project_in (vqgan_vae.py:368-378) feeds the
// codebook arg-max, whose token ids must be the fp32 reference's ids, so its operands may not be rounded
// to bf16.  Every fp32 value is split EXACTLY into three bf16 terms (8 + 8 + 8 mantissa bits),
//   a = a0 + a1 + a2,   a0 = bf16(a), a1 = bf16(a - a0), a2 = bf16(a - a0 - a1),
// and  A W^T = sum over (i + j <= 2) A_i W_j^T  runs as six tcgen05 GEMMs with fp32 accumulation, smallest
// terms first (dropped terms are below 2^-24 relative).  Layout: [rows][3*K] = [x0 | x1 | x2].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, long long ld, bf16* __restrict__ out, long long rows, int K) {
  const long long total = rows * (long long)(K / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (K / 4);
    const int c = (int)(i - r * (K / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    const float a[4] = {v.x, v.y, v.z, v.w};
    uint32_t w[3][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float p0[2], p1[2], p2[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float t = a[2 * h + e];
        p0[e] = __bfloat162float(__float2bfloat16(t));
        const float r1 = t - p0[e];                       // exact (Sterbenz-style: p0 holds the leading bits of t)
        p1[e] = __bfloat162float(__float2bfloat16(r1));
        p2[e] = r1 - p1[e];                               // exact, <= 8 significant bits left
      }
      w[0][h] = pack_bf16x2(p0[0], p0[1]);
      w[1][h] = pack_bf16x2(p1[0], p1[1]);
      w[2][h] = pack_bf16x2(p2[0], p2[1]);
    }
    bf16* o = out + r * 3LL * K + c;
    *reinterpret_cast<uint2*>(o) = make_uint2(w[0][0], w[0][1]);
    *reinterpret_cast<uint2*>(o + K) = make_uint2(w[1][0], w[1][1]);
    *reinterpret_cast<uint2*>(o + 2 * K) = make_uint2(w[2][0], w[2][1]);
  }
}
int split3_f32_bf16(const float* x, long long ld, void* out, long long rows, int K, cudaStream_t stream) {
  if (rows <= 0 || K <= 0 || (K % 4) || (ld % 4) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 7))
    return NUWA_ERR_INVALID;
  const long long g = (rows * (K / 4) + 255) / 256;
  split3_kernel<<<(int)(g > 148LL * 16 ? 148LL * 16 : g), 256, 0, stream>>>(x, ld, reinterpret_cast<bf16*>(out), rows, K);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}
size_t linear_f32x3_workspace(int M, int K) { return (size_t)M * 3 * K * 2; }
// out[M,N] fp32 = x[M,K] fp32 @ W^T + bias, W given pre-split as w3 [N][3K] bf16 (split3_f32_bf16 of the fp32 weight)
int linear_f32x3(const float* x, long long ldx, const void* w3, int M, int N, int K, const float* bias, float* out,
                 void* out_bf16, int ld_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) || out == nullptr || workspace == nullptr) return NUWA_ERR_INVALID;
  if (ws_bytes < linear_f32x3_workspace(M, K)) return NUWA_ERR_WORKSPACE;
  int rc = split3_f32_bf16(x, ldx, workspace, M, K, stream);
  if (rc != NUWA_OK) return rc;
  const bf16* xs = reinterpret_cast<const bf16*>(workspace);
  const bf16* ws = reinterpret_cast<const bf16*>(w3);
  // (i, j): smallest products first; each launch adds the previous partial sum as its fp32 residual (same element,
  // same thread: in place), the last one adds the bias
  static const int order[6][2] = {{2, 0}, {1, 1}, {0, 2}, {1, 0}, {0, 1}, {0, 0}};
  for (int t = 0; t < 6; ++t) {
    const int i = order[t][0], j = order[t][1];
    rc = gemm_bf16(xs + (size_t)i * K, 3 * K, ws + (size_t)j * K, 3 * K, M, N, K, t == 5 ? bias : nullptr,
                   t == 0 ? nullptr : out, ld_out, out, t == 5 ? out_bf16 : nullptr, ld_out, ACT_NONE, 0, stream);
    if (rc != NUWA_OK) return rc;
  }
  return NUWA_OK;
}

// gather rows of an fp32 table -> bf16 (codebook lookup feeding project_out / decode) and/or fp32
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx, bf16* __restrict__ out_bf16,
                   float* __restrict__ out_f32, long long M, int D) {
  const long long total = M * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / D;
    const int d = (int)(i - m * D);
    const float v = table[idx[m] * D + d];
    if (out_bf16) out_bf16[i] = __float2bfloat16(v);
    if (out_f32) out_f32[i] = v;
  }
}
int gather_rows(const float* table, const long long* idx, void* out_bf16, float* out_f32, long long M, int D,
                cudaStream_t stream) {
  if (M <= 0 || D <= 0) return NUWA_ERR_INVALID;
  long long g = (M * D + 255) / 256;
  int grid = (int)(g > 148LL * 32 ? 148LL * 32 : g);
  gather_rows_kernel<<<grid, 256, 0, stream>>>(table, idx, reinterpret_cast<bf16*>(out_bf16), out_f32, M, D);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// final 1x1 convolution dim -> channels (vqgan_vae.py:366), NHWC bf16 in, NCHW fp32 out.
// One warp per pixel: lanes stride the input channels, Cout (<= 8) accumulators, warp-shuffle reduce.
// HBM-bound: reads the activation once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_c1(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// out[b][o][p] = sum_c x[b][p][c] w[o][c] + bias[o],  Cout <= 8, C % 64 == 0.
// One warp = 16 pixels per step: m16n8k16 with the output channels as the (padded) N = 8; the contraction index is
// permuted so that lane t of a quad owns channels [64 blk + 16 t, +16) -> two 16-byte loads per pixel row and block.
// HBM bound: the activation is read exactly once.
__global__ void __launch_bounds__(256)
conv1x1_to_nchw_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ out, long long npix, int HW, int C, int Cout) {
  extern __shared__ __align__(16) uint8_t smem_c1[];
  bf16* ws = reinterpret_cast<bf16*>(smem_c1);  // [8][C] bf16 (rows >= Cout are zero)
  for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) {
    const int o = i / C, c = i - o * C;
    ws[i] = __float2bfloat16(o < Cout ? w[o * C + c] : 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long ntiles = (npix + 15) / 16;
  const int nblk = C / 64;
  for (long long tile = wid; tile < ntiles; tile += nwarps) {
    const long long p0 = tile * 16 + g, p1 = p0 + 8;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int blk = 0; blk < nblk; ++blk) {
      uint4 r0a = make_uint4(0u, 0u, 0u, 0u), r0b = r0a, r1a = r0a, r1b = r0a;
      const int c0 = blk * 64 + 16 * t;
      if (p0 < npix) {
        r0a = __ldg(reinterpret_cast<const uint4*>(x + p0 * C + c0));
        r0b = __ldg(reinterpret_cast<const uint4*>(x + p0 * C + c0) + 1);
      }
      if (p1 < npix) {
        r1a = __ldg(reinterpret_cast<const uint4*>(x + p1 * C + c0));
        r1b = __ldg(reinterpret_cast<const uint4*>(x + p1 * C + c0) + 1);
      }
      const uint32_t a_r0[8] = {r0a.x, r0a.y, r0a.z, r0a.w, r0b.x, r0b.y, r0b.z, r0b.w};
      const uint32_t a_r1[8] = {r1a.x, r1a.y, r1a.z, r1a.w, r1b.x, r1b.y, r1b.z, r1b.w};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // B fragment: output channel g, channels c0 + 4 ks + {0,1 | 2,3}
        const uint2 bw = *reinterpret_cast<const uint2*>(ws + (long long)g * C + c0 + 4 * ks);
        mma_c1(acc, a_r0[2 * ks], a_r1[2 * ks], a_r0[2 * ks + 1], a_r1[2 * ks + 1], bw.x, bw.y);
      }
    }
    // acc[0],acc[1] = (pixel p0, o = 2t, 2t+1) ; acc[2],acc[3] = (pixel p1, ...)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = 2 * t + (e & 1);
      const long long pp = (e < 2) ? p0 : p1;
      if (o < Cout && pp < npix) {
        const long long b = pp / HW;
        out[(b * Cout + o) * HW + (pp - b * HW)] = acc[e] + bias[o];
      }
    }
  }
}
// generic fallback (C % 8 == 0, fp32 weights in shared memory): one warp per pixel
__global__ void __launch_bounds__(256)
conv1x1_to_nchw_scalar_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                              float* __restrict__ out, long long npix, int HW, int C, int Cout) {
  extern __shared__ float w_s[];  // [Cout][C]
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) w_s[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long pix = wid; pix < npix; pix += nwarps) {
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    const bf16* xr = x + pix * C;
    for (int c = lane * 8; c < C; c += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < Cout) {
          const float4 wa = *reinterpret_cast<const float4*>(w_s + o * C + c);
          const float4 wb = *reinterpret_cast<const float4*>(w_s + o * C + c + 4);
          acc[o] += f0.x * wa.x + f0.y * wa.y + f1.x * wa.z + f1.y * wa.w + f2.x * wb.x + f2.y * wb.y + f3.x * wb.z +
                    f3.y * wb.w;
        }
    }
    const long long b = pix / HW;
    const int p = (int)(pix - b * HW);
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < Cout) {
        const float sres = warp_sum(acc[o]);
        if (lane == 0) out[(b * Cout + o) * HW + p] = sres + bias[o];
      }
  }
}
int conv1x1_nhwc_to_nchw(const void* x, const float* w, const float* bias, float* out, int B, int HW, int C, int Cout,
                         cudaStream_t stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || (C % 8) || Cout <= 0 || Cout > 8) return NUWA_ERR_INVALID;
  const long long npix = (long long)B * HW;
  if (C % 64) {
    const size_t smem = (size_t)Cout * C * sizeof(float);
    if (smem > 96 * 1024) return NUWA_ERR_INVALID;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv1x1_to_nchw_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long want = (npix + 7) / 8;
    const int grid = (int)(want > 148LL * 8 ? 148LL * 8 : want);
    conv1x1_to_nchw_scalar_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const bf16*>(x), w, bias, out, npix, HW, C,
                                                                Cout);
    NUWA_CHECK_LAUNCH();
    return NUWA_OK;
  }
  const size_t smem = (size_t)8 * C * sizeof(bf16);
  if (smem > 96 * 1024) return NUWA_ERR_INVALID;
  if (smem > 48 * 1024) cudaFuncSetAttribute(conv1x1_to_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long want = ((npix + 15) / 16 + 7) / 8;
  const int grid = (int)(want > 148LL * 8 ? 148LL * 8 : want);
  conv1x1_to_nchw_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const bf16*>(x), w, bias, out, npix, HW, C, Cout);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
