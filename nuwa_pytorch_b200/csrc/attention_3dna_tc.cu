// Sparse3DNA attention core on tensor cores (causal, 16-wide token grid: 256^2 frames through a 4-layer VAE).
//
// Reference: Sparse3DNA.forward core, nuwa_pytorch.py:523-564 (unfold gather + einsum + mask + softmax +
// talking heads + einsum).  Observation that makes the gather dense: for one ROW of 16 queries (f, y, x=0..15)
// and one (frame, row) offset (a, b) of the (kt, kh, kw) window, the candidate keys are exactly the 16 tokens of
// grid row (f - (kt-1-a) dt, y - (kh-1-b) dh); the window selects the band x' = x - (kw-1-c) dw.  So every
// (a, b) is ONE dense 16x16 score block (mma.sync m16n8k16 over the head dimension) of which kw diagonals are
// used -- 5x redundant tensor work instead of a 46-way scalar gather -- and the whole key set of a query row is
// kt*kh such blocks plus the bos key.  P.V runs block by block the same way through a transposed V copy.
//
// One warp = one query row (16 queries) x all heads (the talking-heads mix needs every head's probabilities of a
// query before any PV): phase 1 scores of all heads -> P[h][x][slot] in shared memory (slot order == the
// reference's key order), fp32 softmax, talking-heads mix, phase 3 PV per head.  Operand fragments are read
// straight from global/L2 with the permuted contraction index of attention_mma.cu (16-byte loads).
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

static constexpr int TC_W = 16;       // grid width handled by this kernel
static constexpr int TC_WARPS = 4;    // query rows per CTA
static constexpr int TC_PITCH = 49;   // floats per P row (odd: conflict-free lane-per-row access), >= 1 + 45 slots
static constexpr int TC_MAXJ = 48;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KS>
__device__ __forceinline__ void ld_words(const bf16* p, uint32_t (&w)[2 * KS]) {
  const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
  w[0] = u0.x; w[1] = u0.y; w[2] = u0.z; w[3] = u0.w;
  if constexpr (KS == 4) {
    const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    w[4] = u1.x; w[5] = u1.y; w[6] = u1.z; w[7] = u1.w;
  }
}

// vT[b][h][d][i] = v of video token i (sequence row 1 + i), zero padded to npad tokens
__global__ void __launch_bounds__(256)
v_transpose_3dna_kernel(const bf16* __restrict__ v, long long v_bs, int v_rs, bf16* __restrict__ vT, int B, int H, int dh,
                        int nv, int npad) {
  __shared__ bf16 tile[32][34];
  // grid: x = token tiles of 32, y = channel tiles of 32 (over H*dh), z = batch
  const int b = blockIdx.z;
  const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    tile[r][tx] = (i < nv) ? v[(long long)b * v_bs + (long long)(1 + i) * v_rs + c0 + tx] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;  // channel = h*dh + d
    const int i = i0 + tx;
    if (i < npad) vT[((long long)b * H * dh + c) * npad + i] = tile[tx][r];
  }
}

template <int DH>
__global__ void __launch_bounds__(TC_WARPS * 32) attn_3dna_tc_kernel(const AttnParams p, const bf16* __restrict__ vT, int npad) {
  constexpr int KS = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int SPAN = DH / 4;
  extern __shared__ float smem_tc[];
  const int H = p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* P = smem_tc + (size_t)warp * H * TC_W * TC_PITCH;  // [H][16][TC_PITCH]
  const int kt = p.kt, kh = p.kh, kw = p.kw;
  const int J = 1 + kt * kh * kw;
  const int rows_per_b = (p.nv + TC_W - 1) / TC_W;  // query rows (of 16 tokens) holding at least one video token
  const int rid = blockIdx.x * TC_WARPS + warp;      // global query-row id
  if (rid >= p.B * rows_per_b) return;
  const int b = rid / rows_per_b;
  const int rr = rid - b * rows_per_b;  // = f*16 + y
  const int f = rr / TC_W, y = rr - f * TC_W;
  const int vbase = rr * TC_W;          // video index of x = 0
  const bf16* qkv_q = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs;
  const bf16* qkv_k = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs;
  const bf16* qkv_v = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs;
  bf16* ob = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs;

  // the bos query (sequence row 0) attends only to itself: copy its value row (done by the first row's warp)
  if (rr == 0)
    for (int c = lane; c < H * DH; c += 32) ob[c] = qkv_v[c];

  // per-lane constants: the 8 (query x, key x') pairs this lane holds in the two 16x8 C tiles of a block, and the
  // window column c they correspond to (or -1)
  int cidx[8];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int x = (e < 2) ? g : g + 8;
      const int xp = nt * 8 + 2 * t + (e & 1);
      const int delta = x - xp;  // causal: key column = x - (kw-1-c)*dw
      int c = -1;
      if (delta >= 0 && delta % p.dw == 0 && delta / p.dw <= kw - 1) c = kw - 1 - delta / p.dw;
      cidx[nt * 4 + e] = c;
    }
  // same for the A fragment of the PV blocks: entry i <-> (x = g + 8*(i&1), key column 4t + (i>>1))
  int aidx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int x = (i & 1) ? g + 8 : g;
    const int delta = x - (4 * t + (i >> 1));
    int c = -1;
    if (delta >= 0 && delta % p.dw == 0 && delta / p.dw <= kw - 1) c = kw - 1 - delta / p.dw;
    aidx[i] = c;
  }
  // (frame,row) blocks of this query row: video index of key column 0, or -1 when the block is outside the grid
  int* blk_tab = reinterpret_cast<int*>(smem_tc + (size_t)TC_WARPS * H * TC_W * TC_PITCH) + warp * 32;
  if (lane < kt * kh) {
    const int a = lane / kh, bq = lane - a * kh;
    const int ff = f - (kt - 1 - a) * p.dt, yy = y - (kh - 1 - bq) * p.dh_;
    blk_tab[lane] = (ff >= 0 && yy >= 0) ? (ff * TC_W + yy) * TC_W : -1;
  }
  const bool ok0 = vbase + g < p.nv, ok1 = vbase + g + 8 < p.nv;  // query validity (partial last row)

  // ---------------- phase 1: scores of every head into P[h][x][slot] ----------------
  for (int i = lane; i < H * TC_W * TC_PITCH; i += 32) P[i] = -FLT_MAX;  // masked unless written
  __syncwarp();  // (also publishes blk_tab)
  for (int h = 0; h < H; ++h) {
    uint32_t qa0[2 * KS], qa1[2 * KS];
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) qa0[i] = qa1[i] = 0u;
    const bf16* qh = qkv_q + h * DH + SPAN * t;
    if (ok0) ld_words<KS>(qh + (long long)(1 + vbase + g) * p.q_rs, qa0);
    if (ok1) ld_words<KS>(qh + (long long)(1 + vbase + g + 8) * p.q_rs, qa1);
    float* Ph = P + (size_t)h * TC_W * TC_PITCH;
    const bf16* kh_ = qkv_k + h * DH + SPAN * t;
    {  // bos key (slot 0): a block whose only non-zero key column is 0
      uint32_t kw_[2 * KS];
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) kw_[i] = 0u;
      if (g == 0) ld_words<KS>(kh_, kw_);  // sequence row 0
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        mma16816(c, qa0[2 * ks], qa1[2 * ks], qa0[2 * ks + 1], qa1[2 * ks + 1], kw_[2 * ks], kw_[2 * ks + 1]);
      if (t == 0) {  // column 0 lives in c[0] (row g) and c[2] (row g+8)
        Ph[g * TC_PITCH] = c[0] * p.qscale;
        Ph[(g + 8) * TC_PITCH] = c[2] * p.qscale;
      }
    }
    // software-pipelined over the kt*kh (frame, row) blocks: the K fragments of block i+1 are in flight while
    // block i is multiplied (the loop is a chain of L2 latencies otherwise)
    const int nblk = kt * kh;
    uint32_t kc0[2 * KS], kc1[2 * KS], kn0[2 * KS], kn1[2 * KS];
    auto load_block = [&](int idx, uint32_t (&d0)[2 * KS], uint32_t (&d1)[2 * KS]) {
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) d0[i] = d1[i] = 0u;
      const int kv0 = idx < nblk ? blk_tab[idx] : -1;
      if (kv0 >= 0) {
        const long long krow = 1 + (long long)kv0;  // sequence row of key x' = 0
        // (only the partial last row of a sequence can have key columns beyond the supplied tokens; they are never
        //  inside a causal window, but they must not be read)
        if (krow + g <= p.nv) ld_words<KS>(kh_ + (krow + g) * p.k_rs, d0);
        if (krow + g + 8 <= p.nv) ld_words<KS>(kh_ + (krow + g + 8) * p.k_rs, d1);
      }
    };
    load_block(0, kc0, kc1);
    for (int idx = 0; idx < nblk; ++idx) {
      load_block(idx + 1, kn0, kn1);
      if (blk_tab[idx] >= 0) {
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          mma16816(c0, qa0[2 * ks], qa1[2 * ks], qa0[2 * ks + 1], qa1[2 * ks + 1], kc0[2 * ks], kc0[2 * ks + 1]);
          mma16816(c1, qa0[2 * ks], qa1[2 * ks], qa0[2 * ks + 1], qa1[2 * ks + 1], kc1[2 * ks], kc1[2 * ks + 1]);
        }
        const int sbase = 1 + idx * kw;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int x = (e < 2) ? g : g + 8;
          if (cidx[e] >= 0) Ph[x * TC_PITCH + sbase + cidx[e]] = c0[e] * p.qscale;
          if (cidx[4 + e] >= 0) Ph[x * TC_PITCH + sbase + cidx[4 + e]] = c1[e] * p.qscale;
        }
      }
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) { kc0[i] = kn0[i]; kc1[i] = kn1[i]; }
    }
  }
  __syncwarp();

  // ---------------- softmax over the J slots of every (head, query) row; lane owns rows lane, lane+32, ... ----------------
  for (int r = lane; r < H * TC_W; r += 32) {
    float* row = P + (size_t)r * TC_PITCH;
    float m = -FLT_MAX;
    for (int j = 0; j < J; ++j) m = fmaxf(m, row[j]);
    float l = 0.f;
    for (int j = 0; j < J; ++j) {
      const float e = __expf(row[j] - m);  // masked slots: exp(-huge) == 0
      row[j] = e;
      l += e;
    }
    const float inv = 1.0f / l;
    for (int j = 0; j < J; ++j) row[j] *= inv;
  }
  __syncwarp();

  // ---------------- talking heads: P'[g][x][j] = sum_h W[g][h] P[h][x][j] ----------------
  if (p.talk != nullptr) {
    const float* Wr = p.talk;  // 64 floats, L1-resident broadcast reads
    for (int i = lane; i < TC_W * J; i += 32) {
      const int x = i / J, j = i - x * J;
      float pin[8];
      for (int h = 0; h < H; ++h) pin[h] = P[((size_t)h * TC_W + x) * TC_PITCH + j];
      for (int gh = 0; gh < H; ++gh) {
        float acc = 0.f;
        for (int h = 0; h < H; ++h) acc = fmaf(__ldg(Wr + gh * H + h), pin[h], acc);
        P[((size_t)gh * TC_W + x) * TC_PITCH + j] = acc;
      }
    }
    __syncwarp();
  }

  // ---------------- phase 3: O_h = sum over blocks P'_blk (16x16) . V_blk (16 keys x DH) ----------------
  for (int h = 0; h < H; ++h) {
    const float* Ph = P + (size_t)h * TC_W * TC_PITCH;
    float o[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    const bf16* vth = vT + ((long long)b * H + h) * DH * npad + 4 * t;
    const int nblk = kt * kh;
    uint2 vc[ND], vn[ND];
    auto load_v = [&](int idx, uint2 (&dst)[ND]) {
      const int kv0 = idx < nblk ? blk_tab[idx] : -1;
      const bool ok = kv0 >= 0;
      const int kbase = ok ? kv0 : 0;  // video index of key x' = 0
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
        dst[nd] = ok ? __ldg(reinterpret_cast<const uint2*>(vth + (long long)(nd * 8 + g) * npad + kbase)) : make_uint2(0u, 0u);
    };
    load_v(0, vc);
    for (int idx = 0; idx < nblk; ++idx) {
      load_v(idx + 1, vn);
      if (blk_tab[idx] >= 0) {
        const int sbase = 1 + idx * kw;
        // A fragment: rows x = g / g+8, contraction slots <-> keys x' = 4t + {0,1 | 2,3}
        float av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int x = (i & 1) ? g + 8 : g;
          av[i] = aidx[i] >= 0 ? Ph[x * TC_PITCH + sbase + aidx[i]] : 0.f;
        }
        // av[2k + half] = P'(x = g + 8*half, key 4t + k)
        const uint32_t a0 = pack_bf16x2(av[0], av[2]);  // row g,   keys 4t, 4t+1
        const uint32_t a1 = pack_bf16x2(av[1], av[3]);  // row g+8, keys 4t, 4t+1
        const uint32_t a2 = pack_bf16x2(av[4], av[6]);  // row g,   keys 4t+2, 4t+3
        const uint32_t a3 = pack_bf16x2(av[5], av[7]);  // row g+8, keys 4t+2, 4t+3
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) mma16816(o[nd], a0, a1, a2, a3, vc[nd].x, vc[nd].y);
      }
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) vc[nd] = vn[nd];
    }
    // bos value (slot 0) + store
    const float pb0 = Ph[g * TC_PITCH], pb1 = Ph[(g + 8) * TC_PITCH];
    const bf16* vbos = qkv_v + h * DH;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int d = nd * 8 + 2 * t;
      const float2 vb = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(vbos + d)));
      if (ok0)
        *reinterpret_cast<uint32_t*>(ob + (long long)(1 + vbase + g) * p.o_rs + h * DH + d) =
            pack_bf16x2(fmaf(pb0, vb.x, o[nd][0]), fmaf(pb0, vb.y, o[nd][1]));
      if (ok1)
        *reinterpret_cast<uint32_t*>(ob + (long long)(1 + vbase + g + 8) * p.o_rs + h * DH + d) =
            pack_bf16x2(fmaf(pb1, vb.x, o[nd][2]), fmaf(pb1, vb.y, o[nd][3]));
    }
  }
}

// Envelope: causal, 16-wide grid, full teacher-forced pass (t0 == 0, nq == 1 + nv), H <= 8, dh in {32, 64},
// window <= 47 keys.  Returns NUWA_ERR_INVALID outside it (the caller then uses the generic kernel).
// vT_ws: B*H*dh*roundup(nv,16) bf16 of scratch.
int attn_3dna_tc(const AttnParams& p, void* vT_ws, cudaStream_t stream) {
  if (!p.causal || p.fmap != TC_W || p.t0 != 0 || p.t0_ptr != nullptr || vT_ws == nullptr) return NUWA_ERR_INVALID;
  if (p.H > 8 || (p.dh != 64 && p.dh != 32) || p.nq != p.nv + 1 || p.nv <= 0) return NUWA_ERR_INVALID;
  if (p.kt * p.kh > 32 || 1 + p.kt * p.kh * p.kw > TC_MAXJ || p.dt <= 0 || p.dh_ <= 0 || p.dw <= 0) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.k_rs % 8) || (p.q_bs % 8) || (p.k_bs % 8) || (p.o_rs & 1)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.k) & 15)) return NUWA_ERR_INVALID;
  const int npad = (p.nv + 15) / 16 * 16;
  bf16* vT = reinterpret_cast<bf16*>(vT_ws);
  dim3 tg((npad + 31) / 32, (p.H * p.dh) / 32, p.B);
  v_transpose_3dna_kernel<<<tg, 256, 0, stream>>>(reinterpret_cast<const bf16*>(p.v), p.v_bs, p.v_rs, vT, p.B, p.H, p.dh,
                                                  p.nv, npad);
  NUWA_CHECK_LAUNCH();
  const int rows = p.B * ((p.nv + TC_W - 1) / TC_W);
  const size_t smem = (size_t)TC_WARPS * p.H * TC_W * TC_PITCH * sizeof(float) + TC_WARPS * 32 * sizeof(int);
  const int grid = (rows + TC_WARPS - 1) / TC_WARPS;
  if (p.dh == 64) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_3dna_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_3dna_tc_kernel<64><<<grid, TC_WARPS * 32, smem, stream>>>(p, vT, npad);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(attn_3dna_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_3dna_tc_kernel<32><<<grid, TC_WARPS * 32, smem, stream>>>(p, vT, npad);
  }
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
