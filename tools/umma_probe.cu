// umma_probe: checks, against a CPU evaluation, the tcgen05 operand conventions the Sparse3DNA tensor-core kernel
// relies on (run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nuwa_pytorch_b200/csrc -o
// gpurun_out/umma_probe tools/umma_probe.cu && gpurun_out/umma_probe):
//   T1  SS mode, both operands K-major SWIZZLE_128B, M=128, N=160 in one instruction; and a sub-range of the same B tile
//       (rows 32..159) written at a TMEM column offset
//   T2  SS mode, B operand MN-major (a [keys][64 ch] V tile exactly as TMA drops it), K = 32 as two K=16 steps
//   T3  TS mode: A operand read from tensor memory (bf16 pairs packed along the columns), B MN-major
//   T4  TS mode with an fp16 A operand and a bf16 B operand in ONE instruction (mixed kind::f16 formats)
//   T5  tcgen05.st / tcgen05.ld round trips of different widths and column offsets
// Exit code 0 and "ALL OK" when every check passes.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "tmem_ldst.cuh"

namespace nuwa { unsigned long long g_launch_count = 0; }
using namespace nuwa;

// byte offset of 16-bit element (row, col) inside a tile of rows of 128 B laid out by TMA with SWIZZLE_128B
__host__ __device__ inline uint32_t sw128(int row, int col) {
  const uint32_t chunk = (uint32_t)(col * 2) >> 4, within = (uint32_t)(col * 2) & 15;
  return (uint32_t)row * 128u + (((chunk ^ (uint32_t)(row & 7))) << 4) + within;
}

struct ProbeArgs {
  const uint16_t* A;   // [128][64] row-major (K contiguous)
  const uint16_t* B;   // test dependent
  float* D;            // [128][256]
  int test;
};

__global__ void __launch_bounds__(128) probe_kernel(const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = sm;                 // 16 KB
  uint8_t* sB = sm + 16384;         // up to 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);

  // A tile: K-major SW128 [128 rows][64]
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    *reinterpret_cast<uint16_t*>(sA + sw128(r, c)) = p.A[i];
  }
  if (p.test == 1) {
    // B: [160 keys][64] K-major
    for (int i = tid; i < 160 * 64; i += 128) *reinterpret_cast<uint16_t*>(sB + sw128(i / 64, i % 64)) = p.B[i];
  } else if (p.test >= 2 && p.test <= 4) {
    // V: [32 keys][64 ch] stored as TMA would (rows = keys, 128 B = 64 channels), used as MN-major B (N = channels)
    for (int i = tid; i < 32 * 64; i += 128) *reinterpret_cast<uint16_t*>(sB + sw128(i / 64, i % 64)) = p.B[i];
  }
  fence_proxy_async_smem();
  __syncthreads();

  if (p.test == 3 || p.test == 4) {
    // A operand -> tensor memory: lane = row, column c holds K elements (2c, 2c+1), low half = even k.  K = 32 -> 16 cols
    uint32_t v[16];
    const int row = warp * 32 + lane;
    for (int c = 0; c < 16; ++c) v[c] = (uint32_t)p.A[row * 64 + 2 * c] | ((uint32_t)p.A[row * 64 + 2 * c + 1] << 16);
    tmem_st_x16(trow + 256, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    if (p.test == 1) {
      const uint32_t idesc160 = make_idesc_f16(128, 160, 1, 1, 0, 0);
      const uint32_t idesc128 = make_idesc_f16(128, 128, 1, 1, 0, 0);
      const uint64_t ad = make_sw128_kmajor_desc(smem_u32(sA)), bd = make_sw128_kmajor_desc(smem_u32(sB));
      const uint64_t bd2 = make_sw128_kmajor_desc(smem_u32(sB) + 2 * 2048);  // rows 32..159
      for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bd + 2 * k, idesc160, k != 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tm + 160 + 32, ad + 2 * k, bd2 + 2 * k, idesc128, k != 0);
    } else if (p.test == 2) {
      const uint32_t idesc = make_idesc_f16(128, 64, 1, 1, 0, 1);  // B MN-major
      const uint64_t ad = make_sw128_kmajor_desc(smem_u32(sA));
      for (int k = 0; k < 2; ++k)
        umma_bf16(tm, ad + 2 * k, make_sw128_kmajor_desc(smem_u32(sB) + k * 2048), idesc, k != 0);
    } else if (p.test == 3 || p.test == 4) {
      const uint32_t idesc = make_idesc_f16(128, 64, p.test == 4 ? 0 : 1, 1, 0, 1);
      for (int k = 0; k < 2; ++k)
        umma_f16_ts(tm, tm + 256 + 8 * k, make_sw128_kmajor_desc(smem_u32(sB) + k * 2048), idesc, k != 0);
    }
    if (p.test != 5) umma_commit(bar);
  }
  if (p.test != 5) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 384; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(trow + c0, v);
      tmem_ld_wait();
      if (c0 < 256) for (int j = 0; j < 32; ++j) p.D[row * 512 + c0 + j] = __uint_as_float(v[j]);
      else for (int j = 0; j < 32; ++j) p.D[row * 512 + c0 + j] = __uint_as_float(v[j]);
    }
  } else {
    // T5: st x8 at col 100 and st x16 at col 37, st x1 at col 7; read back with x64 from 64, x32 from 32, x8 from 0
    const int row = warp * 32 + lane;
    uint32_t z[32];
    for (int j = 0; j < 32; ++j) z[j] = 0;
    for (int c0 = 0; c0 < 160; c0 += 32) tmem_st_x32(trow + c0, z);
    tmem_st_wait();
    uint32_t a8[8], a16[16], a1[1];
    for (int j = 0; j < 8; ++j) a8[j] = 1000u * row + 100 + j;
    for (int j = 0; j < 16; ++j) a16[j] = 1000u * row + 37 + j;
    a1[0] = 1000u * row + 7;
    tmem_st_x8(trow + 100, a8);
    tmem_st_x16(trow + 37, a16);
    tmem_st_x1(trow + 7, a1);
    tmem_st_wait();
    uint32_t r64[64], r32[32], r8[8];
    tmem_ld_x64(trow + 64, r64);
    tmem_ld_x32(trow + 32, r32);
    tmem_ld_x8(trow + 0, r8);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) p.D[row * 512 + j] = (float)r8[j];
    for (int j = 0; j < 32; ++j) p.D[row * 512 + 32 + j] = (float)r32[j];
    for (int j = 0; j < 64; ++j) p.D[row * 512 + 64 + j] = (float)r64[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fff + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2h(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

int main() {
  int ok_all = 1;
  std::vector<uint16_t> A(128 * 64), Ah(128 * 64), B(160 * 64);
  std::vector<float> Af(128 * 64), Ahf(128 * 64), Bf(160 * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) {
    const float v = (rand() % 2001 - 1000) / 500.0f;
    A[i] = f2bf(v); Af[i] = bf2f(A[i]);
    Ah[i] = f2h(v); Ahf[i] = h2f(Ah[i]);
  }
  for (size_t i = 0; i < B.size(); ++i) { B[i] = f2bf((rand() % 2001 - 1000) / 500.0f); Bf[i] = bf2f(B[i]); }
  uint16_t *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 512 * 4);
  std::vector<float> D(128 * 512);
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int test = 1; test <= 5; ++test) {
    cudaMemcpy(dA, test == 4 ? Ah.data() : A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 512 * 4);
    ProbeArgs pa{dA, dB, dD, test};
    probe_kernel<<<1, 128, smem>>>(pa);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T%d: CUDA error %s\n", test, cudaGetErrorString(e)); ok_all = 0; break; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    if (test == 1) {
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 160; ++n) {
        double s = 0; for (int k = 0; k < 64; ++k) s += (double)Af[m * 64 + k] * Bf[n * 64 + k];
        worst = fmax(worst, fabs(s - D[m * 512 + n]));
        if (n >= 32) worst = fmax(worst, fabs(s - D[m * 512 + 160 + n]));  // second product: column = 160 + 32 + (n - 32)
      }
    } else if (test >= 2 && test <= 4) {
      const std::vector<float>& Aa = test == 4 ? Ahf : Af;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double s = 0; for (int k = 0; k < 32; ++k) s += (double)Aa[m * 64 + k] * Bf[k * 64 + n];
        worst = fmax(worst, fabs(s - D[m * 512 + n]));
      }
    } else {
      for (int row = 0; row < 128; ++row) for (int c = 0; c < 128; ++c) {
        float want = 0;
        if (c >= 100 && c < 108) want = 1000.f * row + c;
        if (c >= 37 && c < 53) want = 1000.f * row + c;
        if (c == 7) want = 1000.f * row + 7;
        if (c >= 8 && c < 32) continue;  // not read back
        worst = fmax(worst, fabs(want - D[row * 512 + c]));
      }
    }
    const int ok = worst < 1e-2;
    printf("T%d: max abs err %.3e -> %s\n", test, worst, ok ? "ok" : "FAIL");
    ok_all &= ok;
  }
  printf(ok_all ? "ALL OK\n" : "SOME FAILED\n");
  return ok_all ? 0 : 1;
}
