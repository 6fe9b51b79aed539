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
  const long long total = (long long)B * H * W * Kpad;
  const int pad = KS / 2;
  const int K = KS * KS * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % Kpad);
    const long long pix = i / Kpad;
    float v = 0.f;
    if (kk < K) {
      const int c = kk % C, tap = kk / C;
      const int kh = tap / KS, kw = tap % KS;
      const int x = (int)(pix % W), y = (int)((pix / W) % H);
      const long long b = pix / ((long long)W * H);
      const int yy = y + kh - pad, xx = x + kw - pad;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = img[((b * C + c) * H + yy) * (long long)W + xx];
    }
    out[i] = __float2bfloat16(v);
  }
}
int im2col_nchw_f32(const float* img, void* out, int B, int C, int H, int W, int KS, int Kpad, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || KS <= 0 || Kpad < KS * KS * C || (Kpad % 8)) return NUWA_ERR_INVALID;
  const long long total = (long long)B * H * W * Kpad;
  long long g = (total + 255) / 256;
  int grid = (int)(g > 148LL * 32 ? 148LL * 32 : g);
  im2col_kernel<<<grid, 256, 0, stream>>>(img, reinterpret_cast<bf16*>(out), B, C, H, W, KS, Kpad);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm over NHWC fp32 (vqgan_vae.py:218,221,233,236): stats per (sample, group), then apply
// (+ optional LeakyReLU 0.1) writing bf16 (next conv operand) and/or fp32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gn_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, int HW, int C, int G, float eps) {
  const int b = blockIdx.y, g = blockIdx.x;
  const int cg = C / G;
  const float* base = x + (long long)b * HW * C + g * cg;
  const long long n = (long long)HW * cg;
  // pass 1: mean
  __shared__ double red[16];
  __shared__ float s_mean;
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long p = i / cg;
    const int c = (int)(i - p * cg);
    s += (double)base[p * C + c];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    s_mean = (float)(t / (double)n);
  }
  __syncthreads();
  const float mean = s_mean;
  double q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long p = i / cg;
    const int c = (int)(i - p * cg);
    const float d = base[p * C + c] - mean;
    q += (double)d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    stats[((long long)b * G + g) * 2 + 0] = mean;
    stats[((long long)b * G + g) * 2 + 1] = rsqrtf((float)(t / (double)n) + eps);
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
__global__ void __launch_bounds__(256)
upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C) {
  const int OH = 2 * H, OW = 2 * W;
  const int C8 = C / 8;
  const long long total = (long long)B * OH * OW * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const long long pix = i / C8;
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH);
    const long long b = pix / ((long long)OW * OH);
    float sy = ((float)oy + 0.5f) * 0.5f - 0.5f, sx = ((float)ox + 0.5f) * 0.5f - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const bf16* base = in + b * (long long)H * W * C + c8 * 8;
    const uint4 u00 = *reinterpret_cast<const uint4*>(base + ((long long)y0 * W + x0) * C);
    const uint4 u01 = *reinterpret_cast<const uint4*>(base + ((long long)y0 * W + x1) * C);
    const uint4 u10 = *reinterpret_cast<const uint4*>(base + ((long long)y1 * W + x0) * C);
    const uint4 u11 = *reinterpret_cast<const uint4*>(base + ((long long)y1 * W + x1) * C);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&u00);
    const uint32_t* bq = reinterpret_cast<const uint32_t*>(&u01);
    const uint32_t* cq = reinterpret_cast<const uint32_t*>(&u10);
    const uint32_t* d = reinterpret_cast<const uint32_t*>(&u11);
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f00 = unpack_bf16x2(a[j]), f01 = unpack_bf16x2(bq[j]), f10 = unpack_bf16x2(cq[j]), f11 = unpack_bf16x2(d[j]);
      const float rx = hy * (hx * f00.x + lx * f01.x) + ly * (hx * f10.x + lx * f11.x);
      const float ry = hy * (hx * f00.y + lx * f01.y) + ly * (hx * f10.y + lx * f11.y);
      r[j] = pack_bf16x2(rx, ry);
    }
    *reinterpret_cast<uint4*>(out + pix * C + c8 * 8) = make_uint4(r[0], r[1], r[2], r[3]);
  }
}
int upsample2x_nhwc_bf16(const void* in, void* out, int B, int H, int W, int C, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return NUWA_ERR_INVALID;
  const long long total = (long long)B * 4 * H * W * (C / 8);
  long long g = (total + 255) / 256;
  int grid = (int)(g > 148LL * 32 ? 148LL * 32 : g);
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
static constexpr int VQ_TM = 64, VQ_TN = 64, VQ_TK = 32;

__global__ void __launch_bounds__(256)
vq_argmax_kernel(const float* __restrict__ x, const float* __restrict__ code, const float* __restrict__ code_sq,
                 long long* __restrict__ out, int M, int Kc, int D, int cosine) {
  __shared__ float xs[VQ_TK][VQ_TM + 1];
  __shared__ float cs[VQ_TK][VQ_TN + 1];
  __shared__ float xnorm[VQ_TM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: code sub-tile, ty: token sub-tile
  const int m0 = blockIdx.x * VQ_TM;
  // per-token scale: 1/max(|x|,eps) for cosine, |x|^2 for euclid
  if (tid < VQ_TM) {
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
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += VQ_TK) {
      // load tiles (coalesced along D)
      for (int i = tid; i < VQ_TM * VQ_TK; i += 256) {
        const int r = i / VQ_TK, k = i % VQ_TK;
        const int m = m0 + r;
        float v = (m < M && k0 + k < D) ? x[(long long)m * D + k0 + k] : 0.f;
        if (cosine) v *= xnorm[r];
        xs[k][r] = v;
        const int cn = n0 + r;
        cs[k][r] = (cn < Kc && k0 + k < D) ? code[(long long)cn * D + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < VQ_TK; ++k) {
        float a[4], bq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = xs[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bq[j] = cs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cn = n0 + tx * 4 + j;
        if (cn >= Kc) continue;
        float s = acc[i][j];
        if (!cosine) s = -(xnorm[ty * 4 + i] - 2.0f * s + code_sq[cn]);
        if (s > best[i] || (s == best[i] && cn < besti[i])) { best[i] = s; besti[i] = cn; }
      }
  }
  // arg-max across the 16 lanes (tx) that hold the same tokens: lanes [0,16) and [16,32) of a warp are 2 ty values
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
__global__ void __launch_bounds__(256)
conv1x1_to_nchw_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ out, long long npix, int HW, int C, int Cout) {
  const long long pix = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  const bf16* xr = x + pix * C;
  for (int c = lane * 2; c < C; c += 64) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xr + c));
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < Cout) acc[o] = fmaf(v.x, w[o * C + c], fmaf(v.y, w[o * C + c + 1], acc[o]));
  }
  const long long b = pix / HW;
  const int p = (int)(pix - b * HW);
#pragma unroll
  for (int o = 0; o < 8; ++o)
    if (o < Cout) {
      const float s = warp_sum(acc[o]);
      if (lane == 0) out[(b * Cout + o) * HW + p] = s + bias[o];
    }
}
int conv1x1_nhwc_to_nchw(const void* x, const float* w, const float* bias, float* out, int B, int HW, int C, int Cout,
                         cudaStream_t stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || (C & 1) || Cout <= 0 || Cout > 8) return NUWA_ERR_INVALID;
  const long long npix = (long long)B * HW;
  const int wpb = 8;
  conv1x1_to_nchw_kernel<<<(unsigned)((npix + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      reinterpret_cast<const bf16*>(x), w, bias, out, npix, HW, C, Cout);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
