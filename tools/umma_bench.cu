// umma_bench: issue / execution cost of the small tcgen05.mma shapes the Sparse3DNA kernel uses, one CTA, one issuing
// thread, garbage operands (timing only).  Prints cycles per UMMA for: SS N=160 (QK^T), TS N=64 (P'V, A from TMEM),
// SS N=64, TS N=16 (talking heads); accumulating into ONE accumulator vs alternating between two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nuwa_pytorch_b200/csrc -o tools/umma_bench.bin tools/umma_bench.cu
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "tmem_ldst.cuh"

namespace nuwa { unsigned long long g_launch_count = 0; }
using namespace nuwa;

__global__ void __launch_bounds__(128) bench_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (tid == 0) {
    const uint64_t ad = make_sw128_kmajor_desc(smem_u32(sm)), bd = make_sw128_kmajor_desc(smem_u32(sm) + 16384);
    uint32_t par = 0;
    int o = 0;
    for (int test = 0; test < 8; ++test) {
      const int reps = 64;
      int per = 0;
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t alt = (test & 1) ? (uint32_t)(r & 1) : 0u;  // odd tests alternate between two accumulators
        switch (test >> 1) {
          case 0: {  // SS, N = 160, 4 k-steps
            const uint32_t id = make_idesc_f16(128, 160, 1, 1, 0, 0);
            for (int k = 0; k < 4; ++k) umma_bf16(tm + alt * 160, ad + 2 * k, bd + 2 * k, id, k != 0);
            per = 4;
          } break;
          case 1: {  // TS, N = 64, 10 k-steps, B MN-major
            const uint32_t id = make_idesc_f16(128, 64, 1, 1, 0, 1);
            for (int k = 0; k < 10; ++k) umma_f16_ts(tm + 352 + alt * 64, tm + 8 * k, bd + 128 * k, id, k != 0);
            per = 10;
          } break;
          case 2: {  // SS, N = 64, 10 k-steps (A from smem, K-major), B MN-major
            const uint32_t id = make_idesc_f16(128, 64, 1, 1, 0, 1);
            for (int k = 0; k < 10; ++k) umma_bf16(tm + 352 + alt * 64, ad + 2 * (k & 3), bd + 128 * k, id, k != 0);
            per = 10;
          } break;
          default: {  // TS, N = 16 (talking heads), 2 per slot pair
            const uint32_t id = make_idesc_f16(128, 16, 0, 0, 0, 0);
            for (int k = 0; k < 10; ++k) umma_f16_ts(tm + 192 + 16 * (k >> 1), tm + 8 * (k >> 1), bd, id, k & 1);
            per = 10;
          } break;
        }
      }
      const long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, par);
      par ^= 1;
      const long long t2 = clock64();
      out[o++] = (t1 - t0);
      out[o++] = (t2 - t0);
      out[o++] = reps * per;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048);
  for (int it = 0; it < 2; ++it) bench_kernel<<<1, 128, 65536 + 2048>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[4] = {"SS N=160 K=16 (QK^T)", "TS N=64 K=16 (P'V)", "SS N=64 K=16", "TS N=16 K=16 (mix)"};
  for (int t = 0; t < 8; ++t)
    printf("%-24s %s: issue %.1f cyc/UMMA, issue+drain %.1f cyc/UMMA (%lld UMMAs)\n", names[t >> 1],
           (t & 1) ? "two accumulators" : "one accumulator ", (double)h[3 * t] / h[3 * t + 2], (double)h[3 * t + 1] / h[3 * t + 2],
           h[3 * t + 2]);
  return 0;
}
