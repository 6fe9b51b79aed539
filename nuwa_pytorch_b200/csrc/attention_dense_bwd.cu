// Dense attention backward, probability stage fused: for a tile of 16 queries x 8 heads the kernel RECOMPUTES the logits
// S = Q K^T and dP' = dO V^T on the tensor cores from TMA-staged K / V chunks, keeps the un-normalised probabilities
// (fp16) and dP' (bf16) of the whole tile in shared memory (2 x 72 KB), and then runs the softmax / talking-heads
// backward per query row from there:
//
//   P'[g][j]  = sum_h W[g][h] P[h][j]                      -> HBM (bf16)   (operand of dV = P'^T dO)
//   dP[h][j]  = sum_g W[g][h] dP'[g][j]
//   dS[h][j]  = P[h][j] (dP[h][j] - sum_j' P[h][j'] dP[h][j']) * dh^-0.5   -> HBM (bf16)   (operand of dQ = dS K, dK = dS^T Q)
//   dW[g][h] += sum_j dP'[g][j] P[h][j]
//
// Before: two batched GEMMs wrote S and dP' as fp32 [B][H][nq][jp] (173 MB each per cfg-3 layer) and a row kernel read
// them back (nuwa_attn_bwd_rows); 692 MB of HBM traffic per layer for tensors that only exist between two kernels.
// Differentiates the Attention core of nuwa_pytorch.py:339-378 (decoder text cross-attention, text-encoder self-attention)
// under autograd; the forward formulation (chunk maxima + final factors, exact fp32 null-key logit, key mask as bits)
// is the one of attention_dense_pres.cu, so the recomputed P is the forward's P.
//
// Roles: warps 0-7 = heads in phases 1-2 (S / dP' chunks), = query rows w, w + 8 in phase 3 (16-slot tiles on mma.sync);
// warp 8 = TMA producer (Q tile, dO tile, K half-chunks, V half-chunks through a 4-stage ring of 16 KB).
#include <float.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

int encode_map_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);  // gemm_tcgen05.cu

namespace {

constexpr int BQ = 16;                       // queries per CTA
constexpr int PK = 32;                       // keys per staged chunk
constexpr int NH = 8, DH = 64, INNER = NH * DH;
constexpr int NSTG = 4;                      // 16 KB stages: the ring must keep ~64 KB in flight to cover the L2 latency of one SM's
                                             // TMA feed, and a stage is recycled as soon as its 16 keys sit in registers
constexpr int HK = 16;                       // keys per stage (half a chunk)
constexpr int HBOX = HK * DH * 2;            // 2048 B: one head's [16 keys x 64 channels] box
constexpr int QBOX = BQ * DH * 2;            // 2048 B: one head's [16 queries x 64 channels] box
constexpr int STAGE = NH * HBOX;             // 16 KB (the first two stages hold the Q boxes and the dO boxes of the tile)
constexpr int MAXK = 256;
constexpr int NULLJ = MAXK;                  // slab slot of the null key
constexpr int PP = 280;                      // slab row pitch in 16-bit elements
constexpr int HS = BQ * PP + 8;              // head stride
constexpr int NCF = 12;                      // per (head, query): chunk maxima / final factors (8 chunks + null at 8)

constexpr int OFF_P = NSTG * STAGE;
constexpr int OFF_G = OFF_P + NH * HS * 2;
constexpr int OFF_CF = OFF_G + NH * HS * 2;
constexpr int OFF_SN = OFF_CF + NH * BQ * NCF * 4;
constexpr int OFF_NULL = OFF_SN + NH * BQ * 4;
constexpr int OFF_W = OFF_NULL + 2 * INNER * 4;
constexpr int OFF_DW = OFF_W + NH * NH * 4;
constexpr int OFF_XCH = OFF_DW + NH * NH * 4;          // [4 pairs][2 halves][8 heads]
constexpr int OFF_MASK = OFF_XCH + 4 * 2 * NH * 4;
constexpr int OFF_BAR = OFF_MASK + 64;
constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

struct BwdArgs {
  int B, nq, nk, nchunk, jp, has_null;
  float c1;         // qscale * log2(e)
  float out_scale;  // dS multiplier (dh^-0.5: dQ = dS K and dK = dS^T Q need no further scaling)
  const float* talk;
  float* dtalk;
  const float* null_k;
  const float* null_v;
  const unsigned char* key_mask;
  int mask_bs;
  bf16* Pp;   // [B][H][nq][jp]: slot 0 = null key (when present), slot has_null + j = key j, zero padding up to jp
  bf16* dS;
};

__global__ void __launch_bounds__((NH + 1) * 32, 1)
attn_dense_bwd_fused_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap domap,
                            const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                            const BwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  __half* P16 = reinterpret_cast<__half*>(sm + OFF_P);
  bf16* G16 = reinterpret_cast<bf16*>(sm + OFF_G);
  float* CF = reinterpret_cast<float*>(sm + OFF_CF);      // [h][q][NCF]: chunk maxima, then final factors
  float* SN = reinterpret_cast<float*>(sm + OFF_SN);      // [h][q]: raw null-key logit
  float* nullk = reinterpret_cast<float*>(sm + OFF_NULL);
  float* nullv = nullk + INNER;
  float* Wt = reinterpret_cast<float*>(sm + OFF_W);
  float* dW_cta = reinterpret_cast<float*>(sm + OFF_DW);
  uint32_t* maskw = reinterpret_cast<uint32_t*>(sm + OFF_MASK);  // [8] one bit per key: 1 = attend
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* empty = full + NSTG;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_q = (p.nq + BQ - 1) / BQ;
  const int b = (int)blockIdx.x / tiles_q;
  const int q0 = ((int)blockIdx.x - b * tiles_q) * BQ;
  const int nchunk = p.nchunk;
  const int NS = 2 + 4 * nchunk;  // Q tile, dO tile, then per chunk: two K and two V half-chunks
  const bool has_null = p.has_null != 0;

  // ---- one-time shared state of the 8 consumer warps ----
  auto consumer_state = [&]() {
    for (int i = tid; i < INNER; i += NH * 32) {
      nullk[i] = has_null ? __ldg(p.null_k + i) : 0.f;
      nullv[i] = has_null ? __ldg(p.null_v + i) : 0.f;
    }
    if (tid < NH * NH) {
      Wt[tid] = p.talk != nullptr ? __ldg(p.talk + tid) : ((tid / NH) == (tid % NH) ? 1.f : 0.f);
      dW_cta[tid] = 0.f;
    }
    if (tid < MAXK / 32) {
      const unsigned char* km = p.key_mask != nullptr ? p.key_mask + (long long)b * p.mask_bs : nullptr;
      uint32_t w = 0;
      for (int i = 0; i < 32; ++i) {
        const int j = tid * 32 + i;
        if (j < p.nk && (km == nullptr || km[j] != 0)) w |= 1u << i;
      }
      maskw[tid] = w;
    }
    {  // slab slots of chunks that are never computed, and the null slot when there is no null key: zero in both slabs
      __half* Ph = P16 + (size_t)warp * HS;
      bf16* Gh = G16 + (size_t)warp * HS;
      const int z0 = nchunk * PK, nz = PP - z0;
      for (int i = lane; i < BQ * nz; i += 32) {
        const int q = i / nz, z = i - q * nz;
        Ph[q * PP + z0 + z] = __float2half(0.f);
        Gh[q * PP + z0 + z] = __float2bfloat16(0.f);
      }
    }
  };

  // Producer (warp NH) and consumer warps meet at ONE barrier instruction (single call site, see attention_dense_pres.cu)
  const bool is_producer = warp == NH;
  {
    int st = 0, use = 0;
    auto produce = [&](int s_end) {
      for (int s = use * NSTG + st; s < s_end; ++s) {
        if (use >= 1) mbar_wait(&empty[st], (use - 1) & 1);
        mbar_arrive_expect_tx(&full[st], STAGE);
        if (s < 2) {
          const CUtensorMap* m = s == 0 ? &qmap : &domap;
#pragma unroll
          for (int h = 0; h < NH; ++h) tma_load_3d(sm_u + st * STAGE + h * QBOX, m, &full[st], h * DH, q0, b);
        } else {
          const int hc = s - 2;                       // per 32-key chunk: K half 0, K half 1, V half 0, V half 1
          const int c = hc >> 2, part = hc & 3;
          const CUtensorMap* m = part < 2 ? &kmap : &vmap;
          const int row = c * PK + (part & 1) * HK;
#pragma unroll
          for (int h = 0; h < NH; ++h) tma_load_3d(sm_u + st * STAGE + h * HBOX, m, &full[st], h * DH, row, b);
        }
        if (++st == NSTG) { st = 0; ++use; }
      }
    };
    if (is_producer && lane == 0) {
      for (int i = 0; i < NSTG; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], NH);
      }
      fence_barrier_init();
      tma_prefetch_desc(&qmap);
      tma_prefetch_desc(&domap);
      tma_prefetch_desc(&kmap);
      tma_prefetch_desc(&vmap);
      produce(NSTG);
    }
    if (!is_producer) consumer_state();
    __syncthreads();
    if (is_producer) {
      if (lane == 0) produce(NS);
      return;
    }
  }

  const int h = warp;
  const int mat = lane >> 3, l7 = lane & 7;
  const int k_row = ((mat >> 1) << 3) + l7, k_ch = mat & 1;  // B operand (K, V rows): m0,m1 = keys 0-7 (ch lo, hi); m2,m3 = keys 8-15
  const int a_row = ((mat & 1) << 3) + l7, a_ch = mat >> 1;  // A operand (Q, dO)
  uint32_t k_sw[4], a_sw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    k_sw[i] = k_row * 128 + (((2 * i + k_ch) ^ l7) << 4);
    a_sw[i] = a_row * 128 + (((2 * i + a_ch) ^ l7) << 4);
    asm volatile("" : "+r"(k_sw[i]), "+r"(a_sw[i]));
  }
  __half* Ph = P16 + (size_t)h * HS;
  bf16* Gh = G16 + (size_t)h * HS;
  float* CFh = CF + h * BQ * NCF;
  const float c1 = p.c1;

  int st = 0, par = 0;
  auto release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == NSTG) { st = 0; par ^= 1; }
  };

  // ================= phases 1 + 2 (warp = head): un-normalised probabilities and dP' of head h =================
  {
    uint32_t qa[4][4], da[4][4];
    mbar_wait(&full[0], 0);                       // stages 0 and 1 of the first ring pass: Q boxes, dO boxes
    mbar_wait(&full[1], 0);
    const uint32_t qb = sm_u + h * QBOX;
    const uint32_t db = sm_u + STAGE + h * QBOX;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm4(qa[ks], qb + a_sw[ks]);
      ldsm4(da[ks], db + a_sw[ks]);
    }
    if (has_null && lane < BQ) {  // exact fp32 null-key logit and dP' of the null slot for query `lane`
      float sn = 0.f, dn = 0.f;
      const float* nk_h = nullk + h * DH;
      const float* nv_h = nullv + h * DH;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        uint4 u, w4;
        const uint32_t off = lane * 128 + ((c16 ^ (lane & 7)) << 4);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(qb + off));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w4.x), "=r"(w4.y), "=r"(w4.z), "=r"(w4.w) : "r"(db + off));
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        const float2 e0 = unpack_bf16x2(w4.x), e1 = unpack_bf16x2(w4.y), e2 = unpack_bf16x2(w4.z), e3 = unpack_bf16x2(w4.w);
        const float* w = nk_h + c16 * 8;
        const float* wv = nv_h + c16 * 8;
        sn = fmaf(f0.x, w[0], sn); sn = fmaf(f0.y, w[1], sn); sn = fmaf(f1.x, w[2], sn); sn = fmaf(f1.y, w[3], sn);
        sn = fmaf(f2.x, w[4], sn); sn = fmaf(f2.y, w[5], sn); sn = fmaf(f3.x, w[6], sn); sn = fmaf(f3.y, w[7], sn);
        dn = fmaf(e0.x, wv[0], dn); dn = fmaf(e0.y, wv[1], dn); dn = fmaf(e1.x, wv[2], dn); dn = fmaf(e1.y, wv[3], dn);
        dn = fmaf(e2.x, wv[4], dn); dn = fmaf(e2.y, wv[5], dn); dn = fmaf(e3.x, wv[6], dn); dn = fmaf(e3.y, wv[7], dn);
      }
      SN[h * BQ + lane] = sn;
      Ph[lane * PP + NULLJ] = __float2half(1.0f);  // exp(sn - m) with m = sn as the first running maximum
      Gh[lane * PP + NULLJ] = __float2bfloat16(dn);
      CFh[lane * NCF + 8] = sn;
    }
    release();
    release();

    // row state: r = half <-> query g + 8 * half
    float m_run[2], l_run[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = g + 8 * r;
      m_run[r] = has_null ? SN[h * BQ + q] : -FLT_MAX;
      l_run[r] = (has_null && t == 0) ? 1.f : 0.f;
    }
    for (int c = 0; c < nchunk; ++c) {
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int np = 0; np < 2; ++np) {   // the chunk's two 16-key stages
        mbar_wait(&full[st], par);
        const uint32_t kb = sm_u + st * STAGE + h * HBOX;
        uint32_t kf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm4(kf[ks], kb + k_sw[ks]);
        release();  // the K fragments are in registers
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int n2 = 0; n2 < 2; ++n2) mma_bf16(s[np * 2 + n2], qa[ks], kf[ks][n2 * 2], kf[ks][n2 * 2 + 1]);
      }
      // dP'_h chunk = dO_h V_h^T (V rows are a K-major B operand exactly like K); interleaved with the logits of the same
      // chunk so that the two independent MMA / bookkeeping streams hide each other's latencies
      float d[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        mbar_wait(&full[st], par);
        const uint32_t vb = sm_u + st * STAGE + h * HBOX;
        uint32_t vf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm4(vf[ks], vb + k_sw[ks]);
        release();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int n2 = 0; n2 < 2; ++n2) mma_bf16(d[np * 2 + n2], da[ks], vf[ks][n2 * 2], vf[ks][n2 * 2 + 1]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(Gh + (g + 8 * r) * PP + c * PK + 2 * t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dst[nt * 4] = pack_bf16x2(d[nt][2 * r], d[nt][2 * r + 1]);
      }
      const uint32_t mw = maskw[c];
      if (mw != 0xffffffffu) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (!((mw >> (nt * 8 + 2 * t + e)) & 1u)) s[nt][e] = s[nt][2 + e] = -FLT_MAX;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float cm = -FLT_MAX;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) cm = fmaxf(cm, fmaxf(s[nt][2 * r], s[nt][2 * r + 1]));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
        const float m_new = fmaxf(m_run[r], cm);
        const float mneg = -m_new * c1;
        float l = l_run[r] * fast_exp2(fmaf(m_run[r], c1, mneg));
        const int q = g + 8 * r;
        uint32_t* dst = reinterpret_cast<uint32_t*>(Ph + q * PP + c * PK + 2 * t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = fast_exp2(fmaf(s[nt][2 * r], c1, mneg));      // masked: exp2(-huge) == 0
          const float p1 = fast_exp2(fmaf(s[nt][2 * r + 1], c1, mneg));
          l += p0 + p1;
          dst[nt * 4] = pack_h2(p0, p1);
        }
        m_run[r] = m_new;
        l_run[r] = l;
        if (t == 0) CFh[q * NCF + c] = m_new;
      }
    }
    // final factors: cf[q][c] = exp(m_c - m_final) / sum   (quad lane t handles chunks t, t+4, t+8)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = 1.0f / l, mneg = -m_run[r] * c1;
      const int q = g + 8 * r;
      for (int c = t; c < 9; c += 4) {
        const bool used = c < nchunk || (c == 8 && has_null);
        CFh[q * NCF + c] = used ? fast_exp2(fmaf(CFh[q * NCF + c], c1, mneg)) * inv : 0.f;
      }
    }
  }
  consumer_sync();

  // ================= phase 3 (warp = query rows w and w + 8): softmax / talking-heads backward on the tensor cores =================
  // Per (row, 16-slot tile) -- slots e0 .. e0 + 15 of the slab, 16 key tiles + the null tile -- with mma.sync m16n8k16:
  //   (1) P'[e][g]  = sum_h P16[h][e] (W[g][h] cf[h])     A = slots x heads (fp16, k = 8 .. 15 repeats the heads), B = W cf as
  //                                                        fp16 high | low halves  (the forward kernel's mix)          -> HBM
  //   (2) dP[e][h]  = sum_g dP'[g][e] W[g][h]             A = slots x heads (bf16 dP'), B = W as bf16 high | low
  //   (3) dW[g][h] += sum_e dP'[g][e] P[h][e]             A = dP' rows of head g (slots = contraction), B = normalised P as
  //                                                        bf16 high + low (two MMAs), accumulated in 4 registers per lane
  //   delta[h] = sum_e P[h][e] dP[e][h] from the C fragment of (2): lane (gq, t) holds slots gq, gq + 8 of heads 2t, 2t + 1,
  //   exactly the P values of its A fragment in (1); reduced over gq by shuffles at the end of the row (pass A).
  //   Pass B recomputes (2) per tile and emits dS[h][e] = P (dP - delta) dh^-0.5                                       -> HBM
  {
    const long long hs = (long long)p.nq * p.jp;   // head stride of the outputs
    const uint32_t pa_u = smem_u32(P16) + 2u * (2 * t * HS);    // P16[2t][.][.]; head 2t + 1 is HS elements further
    const uint32_t ga_u = smem_u32(G16) + 2u * (2 * t * HS);
    const uint32_t pr_u = smem_u32(P16) + 2u * (g * HS);        // rows of head g (operands of (3))
    const uint32_t gr_u = smem_u32(G16) + 2u * (g * HS);
    auto lds16 = [](uint32_t addr) -> uint32_t {
      unsigned short v;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
      return v;
    };
    auto lds32 = [](uint32_t addr) -> uint32_t {
      uint32_t v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
      return v;
    };
    auto mma_f16_dup = [](float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%4,%5}, {%6,%7}, {%0,%1,%2,%3};"
          : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
          : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    };
    auto mma_bf16_dup = [](float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%4,%5}, {%6,%7}, {%0,%1,%2,%3};"
          : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
          : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    };
    auto mma_bf16_top = [](float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {   // rows 8 .. 15 of A are zero
      const uint32_t z = 0u;
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%5}, {%7,%8}, {%0,%1,%2,%3};"
          : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
          : "r"(a0), "r"(z), "r"(a2), "r"(b0), "r"(b1));
    };
    auto split_bf16 = [](float x, float y, uint32_t& hi, uint32_t& lo) {
      const bf16 xh = __float2bfloat16(x), yh = __float2bfloat16(y);
      hi = pack_bf16x2(__bfloat162float(xh), __bfloat162float(yh));
      lo = pack_bf16x2(x - __bfloat162float(xh), y - __bfloat162float(yh));
    };
    // B operand of (2): k = source head g' (2t, 2t + 1 | the same + 8 for the low halves), n = gq = destination head h
    const float w2a = Wt[(2 * t) * NH + g], w2b = Wt[(2 * t + 1) * NH + g];
    uint32_t wb_hi, wb_lo;
    split_bf16(w2a, w2b, wb_hi, wb_lo);
    // W[g_out = gq][h = 2t], W[gq][2t + 1] for the B operand of (1), multiplied by the chunk factors per tile
    const float w1a = Wt[g * NH + 2 * t], w1b = Wt[g * NH + 2 * t + 1];
    float dWc[4] = {0.f, 0.f, 0.f, 0.f};   // (3): C[g = gq][h = 2t, 2t + 1] in c0, c1
    constexpr int NT = 17;                 // 16 key tiles + the null tile (slots 256 .. 271)
    for (int q = warp; q < BQ; q += NH) {
      const bool qok = q0 + q < p.nq;
      const long long base = ((long long)b * NH * p.nq + (q0 + q)) * p.jp;
      const float* cfa = CF + ((2 * t) * BQ + q) * NCF;       // factors of heads 2t, 2t + 1 (A-side heads of (1), C heads of (2))
      const float* cfb = cfa + BQ * NCF;
      const float* cfg = CF + (g * BQ + q) * NCF;             // factors of head gq (B operand of (3))
      const uint32_t prow = 2u * (q * PP);
      float d0 = 0.f, d1 = 0.f;                               // partial row sums of heads 2t, 2t + 1
#pragma unroll 1
      for (int tl = 0; tl < NT; ++tl) {
        const int e0 = tl * 16;
        const int ci = tl < 16 ? (tl >> 1) : 8;
        const float f0 = cfa[ci], f1 = cfb[ci], fg = cfg[ci];
        // ---- fragments: slots gq, gq + 8 of heads 2t, 2t + 1 ----
        const uint32_t a = prow + 2u * (e0 + g);
        const uint32_t x00 = lds16(pa_u + a), x01 = lds16(pa_u + 2 * HS + a), x10 = lds16(pa_u + a + 16), x11 = lds16(pa_u + 2 * HS + a + 16);
        const uint32_t y00 = lds16(ga_u + a), y01 = lds16(ga_u + 2 * HS + a), y10 = lds16(ga_u + a + 16), y11 = lds16(ga_u + 2 * HS + a + 16);
        // (1) forward mix
        {
          const float b0f = w1a * f0, b1f = w1b * f1;
          const __half h0 = __float2half_rn(b0f), h1 = __float2half_rn(b1f);
          const __half2 hi = __halves2half2(h0, h1);
          const uint32_t bh = *reinterpret_cast<const uint32_t*>(&hi);
          const uint32_t bl = pack_h2(b0f - __half2float(h0), b1f - __half2float(h1));
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          mma_f16_dup(c, x00 | (x01 << 16), x10 | (x11 << 16), bh, bl);
          if (qok) {   // C[e = gq (+8)][g_out = 2t, 2t + 1]
            const int j0 = e0 + g, j1 = j0 + 8;
            const int o0 = j0 < MAXK ? j0 + p.has_null : ((has_null && j0 == NULLJ) ? 0 : j0);
            const int o1 = j1 < MAXK ? j1 + p.has_null : j1;
            bf16* d = p.Pp + base + (long long)(2 * t) * hs;
            if (o0 < p.jp) { d[o0] = __float2bfloat16(c[0]); d[hs + o0] = __float2bfloat16(c[1]); }
            if (o1 < p.jp) { d[o1] = __float2bfloat16(c[2]); d[hs + o1] = __float2bfloat16(c[3]); }
          }
        }
        // (2) dP and the row sums
        {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16_dup(c, y00 | (y01 << 16), y10 | (y11 << 16), wb_hi, wb_lo);
          const float p00 = __half2float(__ushort_as_half((unsigned short)x00)) * f0, p01 = __half2float(__ushort_as_half((unsigned short)x01)) * f1;
          const float p10 = __half2float(__ushort_as_half((unsigned short)x10)) * f0, p11 = __half2float(__ushort_as_half((unsigned short)x11)) * f1;
          d0 = fmaf(p00, c[0], d0); d0 = fmaf(p10, c[2], d0);
          d1 = fmaf(p01, c[1], d1); d1 = fmaf(p11, c[3], d1);
        }
        // (3) dW: A = dP' rows of head gq over the 16 slots, B = normalised P of head gq (n) ... as [k = slot][n = head]
        {
          const uint32_t r = prow + 2u * (e0 + 2 * t);
          const uint32_t ga0 = lds32(gr_u + r), ga2 = lds32(gr_u + r + 16);        // dP'[gq][e0 + 2t, +1], [e0 + 2t + 8, +9]
          const uint32_t pb0 = lds32(pr_u + r), pb1 = lds32(pr_u + r + 16);        // P16[gq][same slots]
          const __half2 q0h = *reinterpret_cast<const __half2*>(&pb0), q1h = *reinterpret_cast<const __half2*>(&pb1);
          const float2 u0 = __half22float2(q0h), u1 = __half22float2(q1h);
          uint32_t b0h, b0l, b1h, b1l;
          split_bf16(u0.x * fg, u0.y * fg, b0h, b0l);
          split_bf16(u1.x * fg, u1.y * fg, b1h, b1l);
          mma_bf16_top(dWc, ga0, ga2, b0h, b1h);
          mma_bf16_top(dWc, ga0, ga2, b0l, b1l);
        }
      }
      // row sums of heads 2t, 2t + 1 over the slots held by the 8 lanes with this t
      d0 += __shfl_xor_sync(0xffffffffu, d0, 4);  d1 += __shfl_xor_sync(0xffffffffu, d1, 4);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 8);  d1 += __shfl_xor_sync(0xffffffffu, d1, 8);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 16); d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
      if (!qok) continue;
      // ---- pass B: dS ----
#pragma unroll 1
      for (int tl = 0; tl < NT; ++tl) {
        const int e0 = tl * 16;
        const int j0 = e0 + g, j1 = j0 + 8;
        const int o0 = j0 < MAXK ? j0 + p.has_null : ((has_null && j0 == NULLJ) ? 0 : j0);
        const int o1 = j1 < MAXK ? j1 + p.has_null : j1;
        if (__all_sync(0xffffffffu, o0 >= p.jp && o1 >= p.jp)) continue;
        const int ci = tl < 16 ? (tl >> 1) : 8;
        const float f0 = cfa[ci], f1 = cfb[ci];
        const uint32_t a = prow + 2u * (e0 + g);
        const uint32_t x00 = lds16(pa_u + a), x01 = lds16(pa_u + 2 * HS + a), x10 = lds16(pa_u + a + 16), x11 = lds16(pa_u + 2 * HS + a + 16);
        const uint32_t y00 = lds16(ga_u + a), y01 = lds16(ga_u + 2 * HS + a), y10 = lds16(ga_u + a + 16), y11 = lds16(ga_u + 2 * HS + a + 16);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_dup(c, y00 | (y01 << 16), y10 | (y11 << 16), wb_hi, wb_lo);
        const float p00 = __half2float(__ushort_as_half((unsigned short)x00)) * f0, p01 = __half2float(__ushort_as_half((unsigned short)x01)) * f1;
        const float p10 = __half2float(__ushort_as_half((unsigned short)x10)) * f0, p11 = __half2float(__ushort_as_half((unsigned short)x11)) * f1;
        bf16* d = p.dS + base + (long long)(2 * t) * hs;
        if (o0 < p.jp) {
          d[o0] = __float2bfloat16(p00 * (c[0] - d0) * p.out_scale);
          d[hs + o0] = __float2bfloat16(p01 * (c[1] - d1) * p.out_scale);
        }
        if (o1 < p.jp) {
          d[o1] = __float2bfloat16(p10 * (c[2] - d0) * p.out_scale);
          d[hs + o1] = __float2bfloat16(p11 * (c[3] - d1) * p.out_scale);
        }
      }
    }
    if (p.dtalk != nullptr) {
      atomicAdd(&dW_cta[g * NH + 2 * t], dWc[0]);       // C[g = gq][h = 2t], C[gq][2t + 1]
      atomicAdd(&dW_cta[g * NH + 2 * t + 1], dWc[1]);
      consumer_sync();
      if (tid < NH * NH) atomicAdd(p.dtalk + tid, dW_cta[tid]);
    }
  }
}

}  // namespace

// P' and dS of the dense attention backward (see the header of this file).  q: [B][nq] rows (q_bs / q_rs), k / v: [B][nk]
// rows, dO: [B][nq] rows (do_bs / do_rs); Pp / dS: bf16 [B][8][nq][jp] with jp >= nk + has_null, slot 0 = null key.
// Envelope: H == 8, dh == 64, 1 <= nk <= 256; NUWA_ERR_INVALID outside it (nothing launched: the caller keeps the
// materialised-logits path nuwa_bgemm x2 + nuwa_attn_bwd_rows).
int attn_dense_bwd_fused(const AttnParams& p, int nk, const void* dO, long long do_bs, int do_rs, void* Pp, void* dS, int jp,
                         float* dtalk, float out_scale, cudaStream_t stream) {
  if (p.H != NH || p.dh != DH || p.bias != nullptr || p.head_scale != nullptr || p.t0_ptr != nullptr) return NUWA_ERR_INVALID;
  if (nk <= 0 || nk > MAXK || p.nq <= 0 || p.B <= 0 || dO == nullptr || Pp == nullptr || dS == nullptr) return NUWA_ERR_INVALID;
  if ((p.null_k == nullptr) != (p.null_v == nullptr)) return NUWA_ERR_INVALID;
  const int has_null = p.null_k != nullptr;
  if (jp < nk + has_null) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.k_rs % 8) || (p.v_rs % 8) || (p.q_bs % 8) || (p.k_bs % 8) || (p.v_bs % 8) || (do_rs % 8) || (do_bs % 8))
    return NUWA_ERR_INVALID;
  if (p.q_rs < INNER || p.k_rs < INNER || p.v_rs < INNER || do_rs < INNER) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.k) & 15) ||
      (reinterpret_cast<uintptr_t>(p.v) & 15) || (reinterpret_cast<uintptr_t>(dO) & 15))
    return NUWA_ERR_INVALID;

  CUtensorMap qm, dm, km, vm;
  {
    const uint32_t box[3] = {DH, BQ, 1};
    const uint64_t dims[3] = {(uint64_t)INNER, (uint64_t)p.nq, (uint64_t)p.B};
    const uint64_t strq[3] = {2, (uint64_t)p.q_rs * 2, (uint64_t)p.q_bs * 2};
    const uint64_t strd[3] = {2, (uint64_t)do_rs * 2, (uint64_t)do_bs * 2};
    int rc = encode_map_bf16_sw128(&qm, p.q, 3, dims, strq, box);
    if (rc != NUWA_OK) return rc;
    if ((rc = encode_map_bf16_sw128(&dm, dO, 3, dims, strd, box)) != NUWA_OK) return rc;
  }
  {
    const uint32_t box[3] = {DH, HK, 1};
    const uint64_t dims[3] = {(uint64_t)INNER, (uint64_t)nk, (uint64_t)p.B};
    const uint64_t strk[3] = {2, (uint64_t)p.k_rs * 2, (uint64_t)p.k_bs * 2};
    const uint64_t strv[3] = {2, (uint64_t)p.v_rs * 2, (uint64_t)p.v_bs * 2};
    int rc = encode_map_bf16_sw128(&km, p.k, 3, dims, strk, box);
    if (rc != NUWA_OK) return rc;
    if ((rc = encode_map_bf16_sw128(&vm, p.v, 3, dims, strv, box)) != NUWA_OK) return rc;
  }
  BwdArgs a;
  a.B = p.B; a.nq = p.nq; a.nk = nk; a.nchunk = (nk + PK - 1) / PK; a.jp = jp; a.has_null = has_null;
  a.c1 = p.qscale * 1.4426950408889634f;
  a.out_scale = out_scale;
  a.talk = p.talk; a.dtalk = dtalk; a.null_k = p.null_k; a.null_v = p.null_v;
  a.key_mask = p.key_mask; a.mask_bs = p.mask_bs;
  a.Pp = reinterpret_cast<bf16*>(Pp); a.dS = reinterpret_cast<bf16*>(dS);
  static const cudaError_t attr_rc =   // one-time, thread-safe static initialisation, immutable afterwards
      cudaFuncSetAttribute(attn_dense_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr_rc != cudaSuccess) return NUWA_ERR_CUDA;
  const int grid = p.B * ((p.nq + BQ - 1) / BQ);
  attn_dense_bwd_fused_kernel<<<grid, (NH + 1) * 32, SMEM_BYTES, stream>>>(qm, dm, km, vm, a);
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
