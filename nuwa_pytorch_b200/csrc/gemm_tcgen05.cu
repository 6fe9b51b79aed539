// Persistent, warp-specialised bf16 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )        A, W bf16 (K contiguous), fp32 accumulate in TMEM
//
// * operands are staged HBM -> shared memory by TMA (SWIZZLE_128B, 64-element K slabs),
// * the contraction runs on the 5th-gen tensor cores: tcgen05.mma (cta_group::1, M=128, N=BN,
//   K=16 per instruction) issued by one thread, accumulators double-buffered in TMEM so the
//   epilogue of tile i overlaps the main loop of tile i+1,
// * the epilogue (8 warps, two per TMEM lane quadrant so every SM sub-partition has two warps to
//   interleave) reads TMEM with tcgen05.ld, fuses bias, LeakyReLU(0.1), GLU, GEGLU and the residual
//   add, transposes 32x16 slabs through shared memory and writes row-contiguous, fully coalesced
//   fp32 and/or bf16 segments.
//
// "conv" mode turns the A operand into an implicit im2col of an NHWC bf16 activation: one M tile
// is a (tb x th x tw) block of output pixels, each K slab is (filter tap, 64 input channels) and is
// fetched with ONE 4-D TMA box at the tap-shifted coordinate; TMA's out-of-bounds zero fill is the
// convolution's zero padding.  Stride-2 convolutions read one of four parity views of the input
// (one tensor map each), so every tap is still a dense box.
//
// Reference ops this kernel replaces (all ATen library calls in the reference):
//   nn.Linear  nuwa_pytorch.py:274,277,311-313,401-405,1819   nn.Conv2d  vqgan_vae.py:216-238,262-263,352-366
#include <vector>

#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int EPI_WARPS = 8;
static constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;  // warp0 TMA, warp1 MMA(+TMEM alloc), warps2-9 epilogue
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int EPI_PITCH = 20;  // floats per staged row: 16 outputs + 4 pad (conflict-free float4 access)
static constexpr int EPI_SLAB = 4096;  // per epilogue warp: 32 rows x 128 B TMA-store slab (the transposing path uses 2.5 KB of it)
static constexpr int EPI_BYTES = EPI_WARPS * EPI_SLAB;
static constexpr int BAR_BYTES = 1024;  // barrier block; keeps the slabs 1024-byte aligned (SWIZZLE_128B TMA stores)

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator buffers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + EPI_BYTES;
};

struct TileCoord {
  int mt, nt;
};
// grouped raster: walk GROUP_M m-tiles for each n-tile so that concurrently resident CTAs share
// both A and W slabs in L2.
__device__ __forceinline__ TileCoord tile_coord(int tile, int num_m, int num_n) {
  const int GROUP_M = 8;
  int per_group = GROUP_M * num_n;
  int g = tile / per_group;
  int first_m = g * GROUP_M;
  int gsz = min(num_m - first_m, GROUP_M);
  int r = tile - g * per_group;
  TileCoord c;
  c.mt = first_m + (r % gsz);
  c.nt = r / gsz;
  return c;
}

// 16 staged outputs of one accumulator row: bias + activation, written to this thread's staging row.
//   non-pair: outputs = accumulator columns [16*h, 16*h+16) of the 32-column chunk
//   pair    : outputs = value[0,16) combined with gate[16,32)
template <int ACT, int HALF>
__device__ __forceinline__ void epi_stage16(const uint32_t (&v)[32], float* srow, const float* __restrict__ bias,
                                            int ncol /* first accumulator column of the 32-col chunk */, int N) {
#pragma unroll
  for (int j0 = 0; j0 < 16; j0 += 4) {
    float o[4];
    if constexpr (ACT == ACT_GLU || ACT == ACT_GEGLU) {
      float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
      if (bias != nullptr) {
        ba = __ldg(reinterpret_cast<const float4*>(bias + ncol + j0));
        bg = __ldg(reinterpret_cast<const float4*>(bias + ncol + 16 + j0));
      }
      const float a[4] = {__uint_as_float(v[j0]) + ba.x, __uint_as_float(v[j0 + 1]) + ba.y,
                          __uint_as_float(v[j0 + 2]) + ba.z, __uint_as_float(v[j0 + 3]) + ba.w};
      const float g[4] = {__uint_as_float(v[16 + j0]) + bg.x, __uint_as_float(v[17 + j0]) + bg.y,
                          __uint_as_float(v[18 + j0]) + bg.z, __uint_as_float(v[19 + j0]) + bg.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (ACT == ACT_GLU) ? a[j] * sigmoid_f(g[j]) : a[j] * gelu_erf(g[j]);
    } else {
      const int n = ncol + 16 * HALF + j0;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias != nullptr) {
        if (n + 3 < N) {
          bb = __ldg(reinterpret_cast<const float4*>(bias + n));
        } else {
          if (n < N) bb.x = bias[n];
          if (n + 1 < N) bb.y = bias[n + 1];
          if (n + 2 < N) bb.z = bias[n + 2];
        }
      }
      o[0] = __uint_as_float(v[16 * HALF + j0]) + bb.x;
      o[1] = __uint_as_float(v[16 * HALF + j0 + 1]) + bb.y;
      o[2] = __uint_as_float(v[16 * HALF + j0 + 2]) + bb.z;
      o[3] = __uint_as_float(v[16 * HALF + j0 + 3]) + bb.w;
      if constexpr (ACT == ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = leaky01(o[j]);
      }
    }
    *reinterpret_cast<float4*>(srow + j0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Store phase of one 32-row x 16-column slab: lane -> (row = it*8 + lane/4, 4 columns), so a warp store
// instruction covers 8 rows x 64 B (fp32) / 32 B (bf16) of contiguous output.
__device__ __forceinline__ void epi_store16(const float* stg, int lane, const long long (&mrow)[4], unsigned okmask,
                                            int n0, int n_out_total, const GemmParams& p) {
  const int c4 = (lane & 3) * 4;
  const int n = n0 + c4;
  if (n >= n_out_total) return;
  const bool full = n + 3 < n_out_total;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (!((okmask >> it) & 1u)) continue;
    const int rl = it * 8 + (lane >> 2);
    const float4 sv = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + c4);
    float o[4] = {sv.x, sv.y, sv.z, sv.w};
    const long long m = mrow[it];
    if (p.atomic_out) {  // split-K partial product: out += (fp32 reduction in L2)
      if (full) {
        atomicAdd(reinterpret_cast<float4*>(p.out_f32 + m * p.ld_out + n), make_float4(o[0], o[1], o[2], o[3]));
      } else {
        for (int j = 0; j < 4 && n + j < n_out_total; ++j) atomicAdd(p.out_f32 + m * p.ld_out + n + j, o[j]);
      }
      continue;
    }
    if (full) {
      if (p.residual != nullptr) {
        const float4 rr = *reinterpret_cast<const float4*>(p.residual + m * p.ld_res + n);
        o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
      }
      if (p.out_f32 != nullptr)
        *reinterpret_cast<float4*>(p.out_f32 + m * p.ld_out + n) = make_float4(o[0], o[1], o[2], o[3]);
      if (p.out_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o[0], o[1]);
        pk.y = pack_bf16x2(o[2], o[3]);
        *reinterpret_cast<uint2*>(p.out_bf16 + m * p.ld_out + n) = pk;
      }
    } else {
      for (int j = 0; j < 4 && n + j < n_out_total; ++j) {
        float a = o[j];
        if (p.residual != nullptr) a += p.residual[m * p.ld_res + n + j];
        if (p.out_f32 != nullptr) p.out_f32[m * p.ld_out + n + j] = a;
        if (p.out_bf16 != nullptr) p.out_bf16[m * p.ld_out + n + j] = __float2bfloat16(a);
      }
    }
  }
}

template <int ACT>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], float* stg, int lane, const long long (&mrow)[4],
                                          unsigned okmask, int nc, int n_out_total, const GemmParams& p) {
  float* srow = stg + lane * EPI_PITCH;
  if constexpr (ACT == ACT_GLU || ACT == ACT_GEGLU) {
    epi_stage16<ACT, 0>(v, srow, p.bias, nc, p.N);
    __syncwarp();
    epi_store16(stg, lane, mrow, okmask, nc / 2, n_out_total, p);
    __syncwarp();
  } else {
    epi_stage16<ACT, 0>(v, srow, p.bias, nc, p.N);
    __syncwarp();
    epi_store16(stg, lane, mrow, okmask, nc, n_out_total, p);
    __syncwarp();
    if (nc + 16 < p.N) {  // warp-uniform
      epi_stage16<ACT, 1>(v, srow, p.bias, nc, p.N);
      __syncwarp();
      epi_store16(stg, lane, mrow, okmask, nc + 16, n_out_total, p);
      __syncwarp();
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.num_m_tiles * p.num_n_tiles * p.splits;  // split-K: `splits` work items per output tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmC);
    if (p.conv) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmA3);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = p.conv ? (uint32_t)(p.tb * p.th * p.tw * BK * 2) : (uint32_t)A_STAGE_BYTES;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tile = item / p.splits, sp = item - tile * p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        const TileCoord tc = tile_coord(tile, p.num_m_tiles, p.num_n_tiles);
        const int n0 = tc.nt * BN;
        int x0 = 0, y0 = 0, b0 = 0;
        if (p.conv) {
          int xt = tc.mt % p.tiles_x;
          int yt = (tc.mt / p.tiles_x) % p.tiles_y;
          int bt = tc.mt / (p.tiles_x * p.tiles_y);
          x0 = xt * p.tw;
          y0 = yt * p.th;
          b0 = bt * p.tb;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], a_bytes + Cfg::B_STAGE_BYTES);
          if (p.mn_major) {
            // operands stored [contraction][M] / [contraction][N]: one {64 MN elements x 64 contraction rows} box per
            // 64-wide block -> the MN-major SWIZZLE_128B canonical layout, 8 KB per block, blocks consecutive
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &tmA0, &full_bar[stage], tc.mt * BM + i * 64, kb * BK);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &tmB, &full_bar[stage], n0 + i * 64, kb * BK);
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (p.conv) {
            int tap = kb / p.cin_blocks;
            int cb = kb - tap * p.cin_blocks;
            int which = p.tap_map[tap];
            const CUtensorMap* ma = which == 0 ? &tmA0 : (which == 1 ? &tmA1 : (which == 2 ? &tmA2 : &tmA3));
            tma_load_4d(sa, ma, &full_bar[stage], cb * BK, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b0);
          } else {
            tma_load_2d(sa, &tmA0, &full_bar[stage], kb * BK, tc.mt * BM);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // MN-major operands: same instruction, a_major = b_major = 1; descriptors carry the 8 KB pitch between 64-wide
      // blocks as the leading byte offset, and one K = 16 step advances 16 rows of 128 B (2048 B)
      const uint32_t idesc = p.mn_major ? (make_idesc_bf16_f32(BM, BN) | (1u << 15) | (1u << 16)) : make_idesc_bf16_f32(BM, BN);
      const uint64_t kstep = p.mn_major ? 128u : 2u;                       // in 16-byte units
      const uint64_t mn_lbo = p.mn_major ? (((uint64_t)(8192 >> 4) << 16) - ((uint64_t)1 << 16)) : 0;  // replaces LBO = 1
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int sp = item % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(sa) + mn_lbo;
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + A_STAGE_BYTES) + mn_lbo;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in (addr >> 4) units
            umma_bf16(d_tmem, adesc + kstep * k, bdesc + kstep * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue (8 warps) ================================
    const int ew = warp - 2;        // 0..7
    const int q = warp & 3;         // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int half = ew >> 2;       // which half of the tile's column chunks this warp drains
    constexpr int CHUNKS = BN / 32;
    constexpr int CPW = (CHUNKS + 1) / 2;  // chunks per warp
    uint8_t* slab = smem + Cfg::STAGES * Cfg::STAGE_BYTES + BAR_BYTES + ew * EPI_SLAB;
    float* stg = reinterpret_cast<float*>(slab);
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool pair = (p.act == ACT_GLU || p.act == ACT_GEGLU);
    const int n_out_total = pair ? p.N / 2 : p.N;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const TileCoord tc = tile_coord(item / p.splits, p.num_m_tiles, p.num_n_tiles);
      const int n0 = tc.nt * BN;
      // rows this lane stores in the transposed phase: r = q*32 + it*8 + lane/4, it = 0..3
      long long mrow[4];
      unsigned okmask = 0;
      if (p.conv) {
        const int cx0 = (tc.mt % p.tiles_x) * p.tw;
        const int cy0 = ((tc.mt / p.tiles_x) % p.tiles_y) * p.th;
        const int cb0 = (tc.mt / (p.tiles_x * p.tiles_y)) * p.tb;
        const int per_img = p.th * p.tw;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = q * 32 + it * 8 + (lane >> 2);
          const int bb = r / per_img;
          const int rem = r - bb * per_img;
          const int yy = rem / p.tw;
          const int xx = rem - yy * p.tw;
          const int b = cb0 + bb, y = cy0 + yy, x = cx0 + xx;
          if ((bb < p.tb) && (b < p.B) && (y < p.H) && (x < p.W)) okmask |= 1u << it;
          mrow[it] = ((long long)b * p.H + y) * p.W + x;
        }
      } else {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          mrow[it] = (long long)tc.mt * BM + q * 32 + it * 8 + (lane >> 2);
          if (mrow[it] < p.M) okmask |= 1u << it;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int ci = 0; ci < CPW; ++ci) {
        const int c = half * CPW + ci;
        if (c < CHUNKS) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          if (ci == CPW - 1 || c == CHUNKS - 1) {
            // this warp's TMEM reads of the accumulator are done: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          const int nc = n0 + c * 32;
          if (p.tma_out && nc < p.N) {
            // ---- TMA-store epilogue: this lane's accumulator row (32 columns) -> bias / LeakyReLU -> swizzled slab row,
            //      one elected lane hands the 32 x 32 slab to the TMA unit (bounds are clipped by the tensor map; split-K
            //      partial products use the fp32 reduce-add form) ----
            // Pair activations (GLU / GEGLU): a 32-column chunk is 16 value + 16 gate columns -> 16 outputs; the two
            // chunks of an even/odd pair (same warp: CPW is even) fill the left and right half of ONE slab.
            const int sub = pair ? (c & 1) : 0;
            if (sub == 0) {
              if (lane == 0) tma_store_wait_read0();  // the previous store has finished reading the slab
              __syncwarp();
            }
            if (pair) {
              // (warp-uniform branches hoisted out of the element loops: with a branch per element the 16
              //  activations become 16 basic blocks and their dependency chains run one after the other)
              float o[16], g[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                o[j] = __uint_as_float(v[j]);
                g[j] = __uint_as_float(v[16 + j]);
              }
              if (p.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bias + nc + j));
                  const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 16 + j));
                  o[j] += ba.x; o[j + 1] += ba.y; o[j + 2] += ba.z; o[j + 3] += ba.w;
                  g[j] += bg.x; g[j + 1] += bg.y; g[j + 2] += bg.z; g[j + 3] += bg.w;
                }
              }
              if (p.act == ACT_GLU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] *= sigmoid_f(g[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] *= gelu_erf(g[j]);
              }
              if (p.out_f32 != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  *reinterpret_cast<float4*>(slab + lane * 128 + (((sub * 4 + k) ^ (lane & 7)) << 4)) =
                      make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  uint4 u;
                  u.x = pack_bf16x2(o[8 * k], o[8 * k + 1]); u.y = pack_bf16x2(o[8 * k + 2], o[8 * k + 3]);
                  u.z = pack_bf16x2(o[8 * k + 4], o[8 * k + 5]); u.w = pack_bf16x2(o[8 * k + 6], o[8 * k + 7]);
                  *reinterpret_cast<uint4*>(slab + lane * 64 + (((sub * 2 + k) ^ ((lane >> 1) & 3)) << 4)) = u;
                }
              }
              if (sub == 0) continue;  // the odd chunk of the pair completes the slab (N % 64 == 0, host-checked)
            } else {
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) o[j] += __ldg(p.bias + nc + j);
            }
            if (p.act == ACT_LEAKY) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = leaky01(o[j]);
            }
            if (p.out_f32 != nullptr) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(slab + lane * 128 + ((k ^ (lane & 7)) << 4)) =
                    make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint4 u;
                u.x = pack_bf16x2(o[8 * k], o[8 * k + 1]); u.y = pack_bf16x2(o[8 * k + 2], o[8 * k + 3]);
                u.z = pack_bf16x2(o[8 * k + 4], o[8 * k + 5]); u.w = pack_bf16x2(o[8 * k + 6], o[8 * k + 7]);
                *reinterpret_cast<uint4*>(slab + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) = u;
              }
            }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int m0 = tc.mt * BM + q * 32;
              const int ncol = pair ? (nc - 32) / 2 : nc;
              if (p.tma_out == 2) tma_reduce_add_2d(&tmC, slab, ncol, m0);
              else tma_store_2d(&tmC, slab, ncol, m0);
              tma_store_commit();
            }
          } else if (nc < p.N) {  // warp-uniform
            switch (p.act) {
              case ACT_LEAKY: epi_chunk<ACT_LEAKY>(v, stg, lane, mrow, okmask, nc, n_out_total, p); break;
              case ACT_GLU: epi_chunk<ACT_GLU>(v, stg, lane, mrow, okmask, nc, n_out_total, p); break;
              case ACT_GEGLU: epi_chunk<ACT_GEGLU>(v, stg, lane, mrow, okmask, nc, n_out_total, p); break;
              default: epi_chunk<ACT_NONE>(v, stg, lane, mrow, okmask, nc, n_out_total, p); break;
            }
          }
        } else if (ci == CPW - 1) {
          // (only when CHUNKS is odd) nothing to drain in the last slot, still release the accumulator
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.tma_out && lane == 0) tma_store_wait0();  // all bulk stores of this warp are complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

// rank-2..5 bf16 tensor map, innermost box 64 elements (128 B), SWIZZLE_128B, zero OOB fill.
static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return NUWA_ERR_DRIVER;
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  if (rank < 2 || rank > 5) return NUWA_ERR_INVALID;
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NUWA_OK : NUWA_ERR_INVALID;
}

// exported for the other TMA users (attention_3dna_halo.cu)
int encode_map_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_map(map, base, rank, dims, strides_bytes, box);
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static int pick_bn(int num_m_tiles, int N, int force_bn) {
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) return force_bn;
  const int sms = device_sm_count();
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    int bn = cands[i];
    if (bn > 64 && N <= bn / 2) continue;  // too much padding
    int tiles = num_m_tiles * ceil_div(N, bn);
    if (tiles >= sms || bn == 64) return bn;
  }
  return 64;
}

// ---- optional per-launch timing of this kernel (bench.py roofline): CUDA events on the launch stream ----
// No process-global mutable state: a profiler is an object the caller opens, ATTACHES TO ITS HOST THREAD, collects from and
// closes (nuwa_gemm_prof_* in include/nuwa_b200.h).  Launches made by a thread with no profiler attached record nothing;
// two threads with their own profilers do not see each other.
struct GemmProf {
  std::vector<cudaEvent_t> ev;  // start / stop pairs
  size_t used = 0;
  double flops = 0.0, bytes = 0.0, bytes_last = 0.0;
};
static thread_local GemmProf* t_prof = nullptr;
static thread_local double t_cur_flops = 0.0;   // algorithmic work of the call being dispatched on this thread
static thread_local double t_cur_bytes = 0.0;   // algorithmic HBM bytes: operands once + outputs once

void* gemm_prof_open() { return new GemmProf(); }
void gemm_prof_attach(void* h) { t_prof = reinterpret_cast<GemmProf*>(h); }
void gemm_prof_close(void* h) {
  GemmProf* g = reinterpret_cast<GemmProf*>(h);
  if (g == nullptr) return;
  if (t_prof == g) t_prof = nullptr;
  for (cudaEvent_t e : g->ev) cudaEventDestroy(e);
  delete g;
}
double gemm_prof_bytes(void* h) { return h ? reinterpret_cast<GemmProf*>(h)->bytes_last : 0.0; }
// Caller must have synchronised the stream(s).  Returns the number of timed launches and resets the counters.
int gemm_prof_collect(void* h, double* flops, float* ms) {
  GemmProf* g = reinterpret_cast<GemmProf*>(h);
  if (g == nullptr) return 0;
  float total = 0.f;
  for (size_t i = 0; i + 1 < g->used; i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g->ev[i], g->ev[i + 1]) == cudaSuccess) total += t;
  }
  const int n = (int)(g->used / 2);
  if (flops) *flops = g->flops;
  if (ms) *ms = total;
  g->bytes_last = g->bytes;
  g->bytes = 0.0;
  g->flops = 0.0;
  g->used = 0;
  return n;
}
static cudaEvent_t prof_event(GemmProf* g) {
  if (g->used == g->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g->ev.push_back(e);
  }
  return g->ev[g->used++];
}

template <int BN>
static int launch_gemm(const CUtensorMap* mA, const CUtensorMap& mB, const CUtensorMap& mC, GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  // one-time, thread-safe (C++11 static initialisation), immutable afterwards
  static const cudaError_t attr_rc =
      cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (attr_rc != cudaSuccess) return NUWA_ERR_CUDA;
  p.num_n_tiles = ceil_div(p.N, BN);
  if (p.splits <= 1) { p.splits = 1; p.kb_per_split = p.k_blocks; }
  if (p.splits > 1) {  // as many work items as keep every SM busy, every split non-empty
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    int want = ceil_div(2 * device_sm_count(), tiles);
    if (want > p.splits) want = p.splits;
    if (want > p.k_blocks) want = p.k_blocks;
    p.kb_per_split = ceil_div(p.k_blocks, want < 1 ? 1 : want);
    p.splits = ceil_div(p.k_blocks, p.kb_per_split);
  }
  int total = p.num_m_tiles * p.num_n_tiles * p.splits;
  int grid = total < device_sm_count() ? total : device_sm_count();
  if (grid <= 0) return NUWA_OK;
  GemmProf* prof = t_prof;
  if (prof) cudaEventRecord(prof_event(prof), stream);
  gemm_tcgen05_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(mA[0], mA[1], mA[2], mA[3], mB, mC, p);
  if (prof) {
    cudaEventRecord(prof_event(prof), stream);
    prof->flops += t_cur_flops;
    prof->bytes += t_cur_bytes;
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

static int dispatch_gemm(const CUtensorMap* mA, const void* Wp, GemmParams& p, int force_bn, cudaStream_t stream) {
  const int bn = pick_bn(p.num_m_tiles, p.N, force_bn);
  // weight map: dims (K, N), row stride K*2 bytes, box (64, bn)
  CUtensorMap mB;
  uint64_t dimsB[2] = {(uint64_t)p.K, (uint64_t)p.N};
  uint64_t strB[2] = {2, (uint64_t)p.ldw * 2};
  uint32_t boxB[2] = {64, (uint32_t)bn};
  if (p.mn_major) {  // stored [K][N]: the kernel fetches {64 N elements x 64 contraction rows} boxes
    dimsB[0] = (uint64_t)p.N; dimsB[1] = (uint64_t)p.K;
    boxB[1] = 64;
  }
  int e = encode_map(&mB, Wp, 2, dimsB, strB, boxB);
  if (e) return e;
  // Output through TMA stores (32 x 32 slabs) when the epilogue is "bias + optional LeakyReLU" of a plain GEMM with one
  // 16-byte aligned output: removes the per-element address / bounds / transposition instructions that made small-K
  // shapes epilogue-issue bound.  Split-K partial products use the fp32 reduce-add form instead of atomics.
  CUtensorMap mC = mB;
  p.tma_out = 0;
  {
    const bool one_out = (p.out_f32 != nullptr) != (p.out_bf16 != nullptr);
    const bool f32 = p.out_f32 != nullptr;
    const void* outp = f32 ? (const void*)p.out_f32 : (const void*)p.out_bf16;
    const size_t esz = f32 ? 4 : 2;
    const bool pair = (p.act == ACT_GLU || p.act == ACT_GEGLU);
    // pair activations: the two 32-column chunks that share a slab must belong to one epilogue warp (bn >= 128)
    static const bool pair_tma = getenv("NUWA_PAIR_STAGED") == nullptr;  // measurement switch (tools/gemm_epi_perf.py)
    const bool pair_ok = !pair || (pair_tma && bn >= 128 && p.N % 64 == 0 && !p.atomic_out);
    if (!p.conv && p.residual == nullptr && one_out && pair_ok &&
        (reinterpret_cast<uintptr_t>(outp) & 15) == 0 && ((size_t)p.ld_out * esz) % 16 == 0 && (!p.atomic_out || f32)) {
      uint64_t dimsC[2] = {(uint64_t)(pair ? p.N / 2 : p.N), (uint64_t)p.M};
      uint64_t strC[2] = {esz, (uint64_t)p.ld_out * esz};
      uint32_t boxC[2] = {32, 32};
      if (encode_map(&mC, outp, 2, dimsC, strC, boxC, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                     f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B) == NUWA_OK)
        p.tma_out = p.atomic_out ? 2 : 1;
    }
  }
  if (bn == 256) return launch_gemm<256>(mA, mB, mC, p, stream);
  if (bn == 128) return launch_gemm<128>(mA, mB, mC, p, stream);
  return launch_gemm<64>(mA, mB, mC, p, stream);
}

static bool epilogue_args_ok(const GemmParams& p) {
  const bool pair = (p.act == ACT_GLU || p.act == ACT_GEGLU);
  if (p.act < 0 || p.act > ACT_GEGLU) return false;
  if (pair && (p.N % 32) != 0) return false;
  if (p.out_f32 == nullptr && p.out_bf16 == nullptr) return false;
  if (p.ld_out % 4 != 0) return false;
  if (p.residual != nullptr && p.ld_res % 4 != 0) return false;
  if (p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15)) return false;
  return true;
}

int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
              const float* residual, int ld_res, float* out_f32, void* out_bf16, int ld_out, int act, int force_bn,
              cudaStream_t stream, int splits) {
  if (M <= 0 || N <= 0 || K <= 0) return NUWA_ERR_INVALID;
  if (splits > 1 && (bias != nullptr || residual != nullptr || out_bf16 != nullptr || act != ACT_NONE || out_f32 == nullptr))
    return NUWA_ERR_INVALID;  // split-K accumulates raw partial products into out_f32
  if (M <= 32 && force_bn == 0 && splits <= 1) {  // decode-step products: weight-streaming bound, see gemm_skinny.cu
    const int rc = gemm_skinny(A, lda, W, ldw, M, N, K, bias, residual, ld_res, out_f32, out_bf16, ld_out, act, stream);
    if (rc != NUWA_ERR_INVALID) return rc;  // shapes outside its envelope fall through to the tensor-core kernel
  }
  if ((lda % 8) || (ldw % 8) || lda < K || ldw < K) return NUWA_ERR_INVALID;  // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return NUWA_ERR_INVALID;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.ldw = ldw;
  p.k_blocks = ceil_div(K, BK);
  p.num_m_tiles = ceil_div(M, BM);
  p.bias = bias; p.residual = residual; p.ld_res = ld_res;
  p.out_f32 = out_f32; p.out_bf16 = reinterpret_cast<bf16*>(out_bf16); p.ld_out = ld_out; p.act = act;
  p.conv = 0;
  p.splits = splits > 1 ? splits : 1;
  p.atomic_out = splits > 1 ? 1 : 0;  // out_f32 += A W^T even when a single split turns out to be enough
  if (!epilogue_args_ok(p)) return NUWA_ERR_INVALID;
  CUtensorMap mA[4];
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t strA[2] = {2, (uint64_t)lda * 2};
  uint32_t boxA[2] = {64, 128};
  int e = encode_map(&mA[0], A, 2, dimsA, strA, boxA);
  if (e) return e;
  mA[1] = mA[0]; mA[2] = mA[0]; mA[3] = mA[0];
  t_cur_flops = 2.0 * (double)M * (double)N * (double)K;
  {
    const double n_out = (act == ACT_GLU || act == ACT_GEGLU) ? N / 2.0 : (double)N;
    t_cur_bytes = 2.0 * ((double)M * K + (double)N * K) + (double)M * n_out * ((out_f32 ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)) +
                  (residual ? 4.0 * M * n_out : 0.0);
  }
  return dispatch_gemm(mA, W, p, force_bn, stream);
}

// out_f32[M][N] += At^T Wt with At stored [K][M] (row stride lda) and Wt stored [K][N] (row stride ldw), K = the long
// contraction (tokens): the weight-gradient product dW = dY^T X on the activations as the forward / backward passes
// hold them (no transposed copies).  Split-K partial products are reduced with fp32 reduce-adds in L2.
int gemm_bf16_tn_splitk(const void* At, int lda, const void* Wt, int ldw, int M, int N, int K, float* out_f32, int ld_out,
                        int splits, int force_bn, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || out_f32 == nullptr) return NUWA_ERR_INVALID;
  if ((lda % 8) || (ldw % 8) || lda < M || ldw < N) return NUWA_ERR_INVALID;  // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(At) & 15) || (reinterpret_cast<uintptr_t>(Wt) & 15)) return NUWA_ERR_INVALID;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.ldw = ldw;
  p.k_blocks = ceil_div(K, BK);
  p.num_m_tiles = ceil_div(M, BM);
  p.out_f32 = out_f32; p.ld_out = ld_out; p.act = ACT_NONE;
  p.splits = splits > 1 ? splits : 1;
  p.atomic_out = 1;
  p.mn_major = 1;
  if (!epilogue_args_ok(p)) return NUWA_ERR_INVALID;
  CUtensorMap mA[4];
  uint64_t dimsA[2] = {(uint64_t)M, (uint64_t)K};
  uint64_t strA[2] = {2, (uint64_t)lda * 2};
  uint32_t boxA[2] = {64, 64};
  int e = encode_map(&mA[0], At, 2, dimsA, strA, boxA);
  if (e) return e;
  mA[1] = mA[0]; mA[2] = mA[0]; mA[3] = mA[0];
  t_cur_flops = 2.0 * (double)M * (double)N * (double)K;
  t_cur_bytes = 2.0 * ((double)M * K + (double)N * K) + 4.0 * (double)M * N;
  return dispatch_gemm(mA, Wt, p, force_bn, stream);
}

// NHWC bf16 convolution.  x: [B, Hin, Win, Cin] ; w: [Cout, KH*KW, Cin_pad] (Cin_pad = roundup(Cin,64), zero padded)
// supported: (KH=KW=3, stride 1, pad 1), (KH=KW=1, stride 1, pad 0), (KH=KW=4, stride 2, pad 1).
int conv2d_nhwc_bf16(const void* x, const void* w, int B, int Hin, int Win, int Cin, int Cout, int ksize, int stride,
                     const float* bias, const float* residual, float* out_f32, void* out_bf16, int act, int force_bn,
                     cudaStream_t stream) {
  if (B <= 0 || Hin <= 0 || Win <= 0 || Cin <= 0 || Cout <= 0) return NUWA_ERR_INVALID;
  if (Cin % 8) return NUWA_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(x) & 15) return NUWA_ERR_INVALID;
  const bool k3 = (ksize == 3 && stride == 1), k1 = (ksize == 1 && stride == 1), k4 = (ksize == 4 && stride == 2);
  if (!k3 && !k1 && !k4) return NUWA_ERR_INVALID;
  if (k4 && ((Hin & 1) || (Win & 1))) return NUWA_ERR_INVALID;
  const int H = k4 ? Hin / 2 : Hin, W = k4 ? Win / 2 : Win;
  const int cin_blocks = ceil_div(Cin, BK);
  const int ntaps = ksize * ksize;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.conv = 1;
  p.B = B; p.H = H; p.W = W;
  p.tw = W < 128 ? W : 128;
  p.th = (128 / p.tw) < H ? (128 / p.tw) : H;
  p.tb = 128 / (p.tw * p.th);
  if (p.tb > B) p.tb = B;
  if (p.tb < 1) p.tb = 1;
  p.tiles_x = ceil_div(W, p.tw);
  p.tiles_y = ceil_div(H, p.th);
  int tiles_b = ceil_div(B, p.tb);
  p.num_m_tiles = p.tiles_x * p.tiles_y * tiles_b;
  p.cin_blocks = cin_blocks;
  p.M = B * H * W;
  p.N = Cout;
  p.K = ntaps * cin_blocks * BK;
  p.ldw = p.K;
  p.k_blocks = ntaps * cin_blocks;
  p.bias = bias; p.residual = residual;
  const bool pair = (act == ACT_GLU || act == ACT_GEGLU);
  p.ld_res = pair ? Cout / 2 : Cout;
  p.ld_out = pair ? Cout / 2 : Cout;
  p.out_f32 = out_f32; p.out_bf16 = reinterpret_cast<bf16*>(out_bf16); p.act = act;
  if (!epilogue_args_ok(p)) return NUWA_ERR_INVALID;

  CUtensorMap mA[4];
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
  if (!k4) {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    uint64_t str[4] = {2, (uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
    int e = encode_map(&mA[0], xb, 4, dims, str, box);
    if (e) return e;
    mA[1] = mA[0]; mA[2] = mA[0]; mA[3] = mA[0];
    const int pad = k3 ? 1 : 0;
    for (int kh = 0; kh < ksize; ++kh)
      for (int kw = 0; kw < ksize; ++kw) {
        int t = kh * ksize + kw;
        p.tap_map[t] = 0;
        p.tap_dy[t] = (int8_t)(kh - pad);
        p.tap_dx[t] = (int8_t)(kw - pad);
      }
  } else {
    // input row iy = 2*oy + kh - 1 : parity py = (kh-1)&1, half-res row = oy + floor((kh-1)/2)
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)(Win / 2), (uint64_t)(Hin / 2), (uint64_t)B};
        uint64_t str[4] = {2, (uint64_t)2 * Cin * 2, (uint64_t)2 * Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
        uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
        int e = encode_map(&mA[py * 2 + px], xb + ((size_t)py * Win + px) * Cin * 2, 4, dims, str, box);
        if (e) return e;
      }
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        int t = kh * 4 + kw;
        int py = (kh - 1) & 1, px = (kw - 1) & 1;
        p.tap_map[t] = (int8_t)(py * 2 + px);
        p.tap_dy[t] = (int8_t)((kh - 1 - py) / 2);
        p.tap_dx[t] = (int8_t)((kw - 1 - px) / 2);
      }
  }
  t_cur_flops = 2.0 * (double)p.M * (double)Cout * (double)(ntaps * Cin);  // algorithmic (un-padded) MACs x 2
  {
    const double n_out = (act == ACT_GLU || act == ACT_GEGLU) ? Cout / 2.0 : (double)Cout;
    t_cur_bytes = 2.0 * ((double)B * Hin * Win * Cin + (double)Cout * ntaps * Cin) +
                  (double)p.M * n_out * ((out_f32 ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)) + (residual ? 4.0 * p.M * n_out : 0.0);
  }
  return dispatch_gemm(mA, w, p, force_bn, stream);
}

}  // namespace nuwa
