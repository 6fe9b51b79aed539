// tma_bw: how fast can every SM pull Sparse3DNA key tiles (10 boxes of [16 tokens x 64 channels] of a q|k|v buffer with
// 3072-byte token rows, as 8-row + 2-row 5-D TMA boxes) out of L2 / HBM?  One producer thread per CTA, ring of NST
// stages, no compute.  Prints aggregate GB/s for a 63 MB buffer (L2 resident after the first pass) and per-SM B/clk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nuwa_pytorch_b200/csrc -o tools/tma_bw.bin tools/tma_bw.cu
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace nuwa { unsigned long long g_launch_count = 0; }
using namespace nuwa;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int NST>
__global__ void __launch_bounds__(64) bw_kernel(const __grid_constant__ CUtensorMap m8, const __grid_constant__ CUtensorMap m2,
                                                 int tiles_per_cta, int nrow_blocks, int B, long long* cycles, int head_major) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NST * 20480);
  uint64_t* empty = full + NST;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < tiles_per_cta; ++it) {
      const int st = it % NST;
      mbar_wait(&empty[st], ((it / NST) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[st], 20480);
      // pseudo-random tile: (sample, head, k|v, row block)
      const unsigned u = (unsigned)(blockIdx.x * 7919 + it * 104729);
      const int b = u % B, h = (u / 7) % 8, kv = 1 + (u / 3) % 2, rb = (u / 11) % nrow_blocks;
      const int c0 = head_major ? 0 : kv * 512 + h * 64, c4 = head_major ? b * 24 + kv * 8 + h : b;
      if (head_major == 2) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sm) + st * 20480), "l"(reinterpret_cast<uint64_t>(&m8)), "r"(smem_u32(&full[st])), "r"(0),
                       "r"((c4 * 160 + rb) * 16) : "memory");
        continue;
      }
      tma5(smem_u32(sm) + st * 20480, &m8, &full[st], c0, 0, 0, rb, c4);
      tma5(smem_u32(sm) + st * 20480 + 16384, &m2, &full[st], c0, 0, 0, rb + 8, c4);
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < tiles_per_cta; ++it) {
      const int st = it % NST;
      mbar_wait(&full[st], (it / NST) & 1);
      mbar_arrive(&empty[st]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

static int g_head_major = 0;

int main() {
  const int B = 8, NTOK = 2560, ROW = 1536;
  const size_t bytes = (size_t)B * NTOK * ROW * 2;
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)f;
  CUtensorMap m8, m2;
  const cuuint64_t dims[5] = {ROW, 16, 1, (cuuint64_t)(NTOK / 16), B};
  const cuuint64_t strides[4] = {(cuuint64_t)ROW * 2, (cuuint64_t)16 * ROW * 2, (cuuint64_t)16 * ROW * 2, (cuuint64_t)NTOK * ROW * 2};
  const cuuint32_t b8[5] = {64, 16, 1, 8, 1}, b2[5] = {64, 16, 1, 2, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult r1 = enc(&m8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, b8, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int tiles = 400, nrb = NTOK / 16 - 10;
  auto run = [&](auto kern, int nst, const char* name) {
    const int smem = nst * 20480 + 256 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      kern<<<148, 64, smem>>>(m8, m2, tiles, nrb, B, cyc, g_head_major);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double total = 148.0 * tiles * 20480;
    printf("%s: %.1f us, %.2f TB/s aggregate, %.1f B/clk/SM (%.0f cycles per 20 KB tile per SM)\n", name, ms * 1e3,
           total / (ms * 1e-3) / 1e12, tiles * 20480.0 / avg, avg / tiles);
  };
  run(bw_kernel<2>, 2, "ring of 2 stages");
  run(bw_kernel<4>, 4, "ring of 4 stages");
  run(bw_kernel<6>, 6, "ring of 6 stages");
  run(bw_kernel<10>, 10, "ring of 10 stages");
  // ---- the same tiles out of a HEAD-MAJOR buffer [B][24 (q|k|v x head)][token][64]: a grid row of 16 tokens is 2 KB
  //      contiguous (token pitch 128 B instead of 3072 B); coordinates: c0 = channel (0), c4 = b * 24 + kv * 8 + h ----
  {
    const cuuint64_t dimsH[5] = {64, 16, 1, (cuuint64_t)(NTOK / 16), (cuuint64_t)B * 24};
    const cuuint64_t stridesH[4] = {128, 2048, 2048, (cuuint64_t)NTOK * 128};
    r1 = enc(&m8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dimsH, stridesH, b8, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r2 = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dimsH, stridesH, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode (head-major) failed %d %d\n", (int)r1, (int)r2); return 1; }
    g_head_major = 1;
    run(bw_kernel<2>, 2, "head-major, ring of 2 stages");
    run(bw_kernel<4>, 4, "head-major, ring of 4 stages");
    run(bw_kernel<10>, 10, "head-major, ring of 10 stages");
  }
  // ---- the same bytes as ONE 2-D box {64 channels, 160 token rows} of the head-major buffer (dilation-1 tiles are
  //      contiguous there): is the 5-D box walk the limiter? ----
  {
    const cuuint64_t dims2[2] = {64, (cuuint64_t)NTOK * 24 * B};
    const cuuint64_t strides2[1] = {128};
    const cuuint32_t bx[2] = {64, 160}, es2[2] = {1, 1};
    r1 = enc(&m8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims2, strides2, bx, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS) { printf("encode (2-D) failed %d\n", (int)r1); return 1; }
    g_head_major = 2;
    run(bw_kernel<2>, 2, "head-major 2-D box 64 x 160, ring of 2 stages");
    run(bw_kernel<4>, 4, "head-major 2-D box 64 x 160, ring of 4 stages");
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
