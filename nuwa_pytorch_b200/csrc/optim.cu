// Trainer step over the flat fp32 buffers of the training path: global gradient norm, clip, AdamW, zero_grad.
//
// Replaces NUWATrainer.train_step's tail (train_nuwa.py:256-258: torch.nn.utils.clip_grad_norm_(params, max_norm);
// optim.step(); optim.zero_grad()) with the optimizer of optimizer.py:11-31 (AdamW, weight decay only on
// parameters with ndim >= 2).  The reference runs ~10 ATen kernels per parameter tensor (foreach or not); here the
// parameters, gradients and both moments are four flat fp32 buffers with ONE layout (train.GradStore), so the step is
// two HBM-bound passes:
//   sqnorm:  sum of squares of the gradient buffer, two deterministic stages (per-CTA partials in fixed order, then
//            one CTA) -- no float atomics, identical on every data-parallel rank,
//   adamw:   clip coefficient from the device-side norm (no host sync), decoupled weight decay, moment updates, bias
//            correction, parameter update and the zeroing of the gradient for the next accumulation, in one pass:
//            read p, g, m, v (16 B / element), write p, m, v, g (16 B / element).
// Per-parameter attributes (weight decay on / off) come from a table of fixed-size chunks, so that one launch covers
// every parameter tensor (multi-tensor apply without per-tensor launches).
#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int OPT_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (warp == 0) {
    s = lane < OPT_THREADS / 32 ? red[lane] : 0.f;
    s = warp_sum(s);
  }
  __syncthreads();
  return s;  // valid in warp 0
}

__global__ void __launch_bounds__(OPT_THREADS) sqnorm_partial_kernel(const float* __restrict__ x, long long n,
                                                                      float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  const long long n4 = n >> 2;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * OPT_THREADS) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(OPT_THREADS) sqnorm_final_kernel(const float* __restrict__ partials, int nparts,
                                                                    float* __restrict__ out, int accumulate) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += OPT_THREADS) acc += partials[i];
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
}

int sqnorm_f32(const float* x, long long n, float* partials, int nparts, float* out, int accumulate, cudaStream_t stream) {
  if (x == nullptr || partials == nullptr || out == nullptr || n <= 0 || nparts <= 0) return NUWA_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(x) & 15) return NUWA_ERR_INVALID;
  sqnorm_partial_kernel<<<nparts, OPT_THREADS, 0, stream>>>(x, n, partials);
  NUWA_CHECK_LAUNCH();
  sqnorm_final_kernel<<<1, OPT_THREADS, 0, stream>>>(partials, nparts, out, accumulate);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// mean |a - b| (l1) or mean (a - b)^2 (l2) of two fp32 arrays: VQGanVAE.forward(return_loss=True) reconstruction loss
// without the GAN / perceptual terms (vqgan_vae.py:340, :502-512).  Same two deterministic stages as sqnorm.
__global__ void __launch_bounds__(OPT_THREADS) diff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                    long long n, int l2, float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc += l2 ? d * d : fabsf(d);
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(OPT_THREADS) mean_final_kernel(const float* __restrict__ partials, int nparts,
                                                                  float inv_n, float* __restrict__ out) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += OPT_THREADS) acc += partials[i];
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) out[0] = s * inv_n;
}
int recon_loss_f32(const float* a, const float* b, long long n, int l2, float* partials, int nparts, float* out,
                   cudaStream_t stream) {
  if (a == nullptr || b == nullptr || partials == nullptr || out == nullptr || n <= 0 || nparts <= 0) return NUWA_ERR_INVALID;
  diff_partial_kernel<<<nparts, OPT_THREADS, 0, stream>>>(a, b, n, l2, partials);
  NUWA_CHECK_LAUNCH();
  mean_final_kernel<<<1, OPT_THREADS, 0, stream>>>(partials, nparts, 1.0f / (float)n, out);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

// one CTA per chunk; a chunk never crosses a parameter boundary
__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const nuwa_adamw_params a) {
  const nuwa_opt_chunk ck = a.chunks[blockIdx.x];
  float clip = 1.0f;
  if (a.sqnorm != nullptr && a.max_norm > 0.f) {
    const float total = sqrtf(__ldg(a.sqnorm)) * a.grad_scale;   // norm of the scaled gradient
    clip = fminf(a.max_norm / (total + 1e-6f), 1.0f);             // torch.nn.utils.clip_grad_norm_
  }
  const float gs = a.grad_scale * clip;
  const int step = a.step_ptr != nullptr ? __ldg(a.step_ptr) : a.step;
  const float bc1 = 1.0f - powf(a.beta1, (float)step), bc2 = 1.0f - powf(a.beta2, (float)step);
  const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = ck.weight_decay ? 1.0f - a.lr * a.weight_decay : 1.0f;
  float* __restrict__ p = a.p + ck.offset;
  float* __restrict__ g = a.g + ck.offset;
  float* __restrict__ m = a.m + ck.offset;
  float* __restrict__ v = a.v + ck.offset;
  const int n4 = ck.len >> 2;  // offsets are multiples of 4 elements (16-byte aligned rows of the flat layout)
  for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * gs;
      pp[k] *= decay;
      mp[k] = a.beta1 * mp[k] + (1.0f - a.beta1) * gk;
      vp[k] = a.beta2 * vp[k] + (1.0f - a.beta2) * gk * gk;
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + a.eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (a.zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < ck.len; i += OPT_THREADS) {
    const float gk = g[i] * gs;
    float pk = p[i] * decay;
    const float mk = a.beta1 * m[i] + (1.0f - a.beta1) * gk;
    const float vk = a.beta2 * v[i] + (1.0f - a.beta2) * gk * gk;
    pk -= step_size * (mk / (sqrtf(vk) * inv_sqrt_bc2 + a.eps));
    p[i] = pk; m[i] = mk; v[i] = vk;
    if (a.zero_grad) g[i] = 0.f;
  }
}

int adamw_step(const nuwa_adamw_params& a, cudaStream_t stream) {
  if (a.p == nullptr || a.g == nullptr || a.m == nullptr || a.v == nullptr || a.chunks == nullptr || a.nchunks <= 0)
    return NUWA_ERR_INVALID;
  if (a.step_ptr == nullptr && a.step < 1) return NUWA_ERR_INVALID;
  adamw_kernel<<<a.nchunks, OPT_THREADS, 0, stream>>>(a);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
