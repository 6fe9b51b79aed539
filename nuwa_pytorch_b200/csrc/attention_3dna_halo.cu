// Sparse3DNA attention core, halo-tiled: K/V key rows staged in shared memory by TMA, dense banded 16x16 score /
// PV blocks on the tensor cores (causal, 16-wide token grid, 8 heads x 64).
//
// Reference: Sparse3DNA.forward core, nuwa_pytorch.py:523-564 (unfoldNd gather of k,v -> einsum -> mask -> softmax
// -> talking heads -> einsum).  The reference (and attention.cu's gather kernel) reads 46 key rows PER QUERY;
// here one CTA owns QR = 4 query rows of a frame, (f, y0 + i*dil_h, x = 0..15), chosen with the window's own row
// dilation so that their key rows overlap: for every frame offset a of the window the CTA needs only the
// QR + kh - 1 = 6 key rows (f_a, y0 + (r - (kh-1))*dil_h), each a [16 tokens x 64 channels] box per head that one
// TMA instruction (SWIZZLE_128B tensor map over the q|k|v projection buffer, zero fill beyond the sequence) drops
// into a 3-stage shared-memory ring -- 7.5 key rows per query row instead of 15 (row-per-warp) or 46 x 16 (gather).
//
// One warp = one query row (16 queries) and walks the heads one after the other:
//   phase 1, head h: Q fragments (ldmatrix) x K row blocks (ldmatrix) -> mma.sync m16n8k16 -> the kw in-band
//            diagonals of each 16x16 block go to an fp32 row buffer S[x][slot] (slot order = the reference's key
//            order, slot 0 = bos); fp32 softmax -> P[h][q][slot] as fp16;
//   mix:     talking heads across the 8 heads per (query, slot), fp32 math, result stored as bf16 in place;
//   phase 3, head g: banded P'[g] blocks (A fragments gathered from P) x V row blocks (ldmatrix.trans) -> O.
// Out-of-grid window slots keep the -FLT_MAX the row buffer is initialised with (causal: no zero-key case, D16
// cannot occur).  Everything except the stage ring is warp-private; a fifth warp is the TMA producer and the ring
// is handed over through full / empty mbarriers, so the query-row warps never wait for each other.
#include <float.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace nuwa {

int encode_map_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);  // gemm_tcgen05.cu

namespace {

constexpr int GW = 16;                      // tokens per grid row
constexpr int QR = 4;                       // query rows (warps) per CTA
constexpr int NR = 6;                       // key rows per frame offset: QR + kh - 1 with kh <= 3
constexpr int NST = 3;                      // stage ring depth
constexpr int NH = 8, DH = 64, INNER = NH * DH;
constexpr int BOX = GW * DH * 2;            // 2048 B: one (row, head) box
constexpr int STAGE = NR * BOX;
constexpr int SP = 49;                      // fp32 score row pitch
constexpr int PP = 50;                      // fp16/bf16 probability row pitch (25 words: conflict-free across rows)
constexpr int ZSLOT = PP - 1;               // always-zero slot
constexpr int MAXJ = 48;

constexpr int OFF_Q = NST * STAGE;
constexpr int OFF_P = OFF_Q + QR * BOX;
constexpr int OFF_S = OFF_P + NH * QR * GW * PP * 2;
constexpr int OFF_BOS = OFF_S + QR * GW * SP * 4;
constexpr int OFF_W = OFF_BOS + 2 * INNER * 2;
constexpr int OFF_TRASH = OFF_W + NH * NH * 4;     // one word per query-row lane: out-of-band score entries land here
constexpr int OFF_BAR = OFF_TRASH + QR * 32 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // + alignment slack

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
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

// The producer warp and the query-row warps meet at ONE CTA-wide barrier, at a single call site.
__device__ __forceinline__ void cta_bar() { __syncthreads(); }

struct HaloArgs {
  int B, nv, nf, tiles;
  int kt, kh, kw, dt, dh, dw;
  int koff, voff;  // channel offsets of k and v relative to q inside a token row
  int J;
  float scale_log2e;
  const float* talk;
  bf16* o;
  long long o_bs;
  int o_rs;
  const bf16* k0;  // k of sequence row 0 (bos), sample 0
  const bf16* v0;
  long long k_bs, v_bs;
  long long* dbg;  // tools/halo_stamps.py: clock64 stamps of CTA 0 / warp 0 (NULL in normal use)
};

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

// Warps 0..QR-1: one query row each.  Warp QR: TMA producer (one lane).  Stage hand-over through full / empty
// mbarriers, so the query-row warps never wait for each other.
template <int KH, bool DBG>
__global__ void __launch_bounds__((QR + 1) * 32, 2)
attn_3dna_halo_kernel(const __grid_constant__ CUtensorMap qmap, const HaloArgs p) {
  constexpr int NROW = QR + KH - 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  __half* P16 = reinterpret_cast<__half*>(sm + OFF_P);
  float* S32 = reinterpret_cast<float*>(sm + OFF_S);
  bf16* kbos = reinterpret_cast<bf16*>(sm + OFF_BOS);
  bf16* vbos = kbos + INNER;
  float* Wsm = reinterpret_cast<float*>(sm + OFF_W);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);  // [NST]
  uint64_t* empty = full + NST;                                // [NST]
  uint64_t* qfull = empty + NST;
  uint64_t* qempty = qfull + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long t_start = clock64();

  // ---- which tile: heaviest frames first (later frames see more key frames) ----
  const int per_f = p.tiles * p.B;
  const int f = p.nf - 1 - (int)blockIdx.x / per_f;
  const int rem = (int)blockIdx.x % per_f;
  const int tile = rem / p.B, b = rem - tile * p.B;
  int y0;
  {
    int r = 0, s = tile;
    for (; r < p.dh; ++r) {
      const int nrows = (GW - r + p.dh - 1) / p.dh;
      const int nt = (nrows + QR - 1) / QR;
      if (s < nt) break;
      s -= nt;
    }
    y0 = r + s * QR * p.dh;
  }
  const int kt = p.kt, kw = p.kw;
  const int a_lo = max(0, kt - 1 - f / p.dt);
  const int n_a = kt - a_lo;
  const int NS1 = NH * n_a, NS = 2 * NS1;

  // ---- one-time shared-memory state of the query-row warps ----
  auto one_time_state = [&]() {
    float* Sw = S32 + warp * GW * SP;
    for (int i = lane; i < GW * SP; i += 32) Sw[i] = -FLT_MAX;
    const int nz = PP - p.J;  // slots J .. PP-1 stay zero (pair padding + the always-zero slot)
    for (int i = lane; i < NH * GW * nz; i += 32) {
      const int row = i / nz, z = i - row * nz;
      const int h = row / GW, x = row - h * GW;
      P16[(size_t)(h * QR * GW + warp * GW + x) * PP + p.J + z] = __float2half(0.f);
    }
    // bos key / value rows of this sample (sequence row 0), all heads
    const uint4* src = reinterpret_cast<const uint4*>(tid < 64 ? p.k0 + (long long)b * p.k_bs : p.v0 + (long long)b * p.v_bs);
    reinterpret_cast<uint4*>(kbos)[tid] = __ldg(src + (tid & 63));
    if (tid < NH * NH) Wsm[tid] = p.talk ? __ldg(p.talk + tid) : ((tid / NH) == (tid % NH) ? 1.f : 0.f);
  };

  // =============================== producer state (warp QR) ===============================
  // Lane 0 initialises the barriers and starts the first loads BEFORE the CTA-wide sync, so the first key rows
  // (DRAM latency: the layer's q|k|v was just written by the projection GEMM) travel while the query-row warps
  // set up their shared-memory state.  Both roles meet at ONE barrier instruction (a single call site: the two
  // roles arriving from different program counters is what compute-sanitizer synccheck reports as divergence).
  const bool is_producer = warp == QR;
  {
    uint32_t qmask = 0, rowmask = 0;
    int tok_q[QR], tok_k[NROW];
#pragma unroll
    for (int i = 0; i < QR; ++i) {
      const int y = y0 + i * p.dh;
      if (y < GW) qmask |= 1u << i;
      tok_q[i] = 1 + (f * GW + y) * GW;
    }
#pragma unroll
    for (int rr = 0; rr < NROW; ++rr) {
      const int yy = y0 + (rr - (KH - 1)) * p.dh;
      if (yy >= 0 && yy < GW) rowmask |= 1u << rr;
      tok_k[rr] = 1 + yy * GW;
    }
    const uint32_t qbytes = (uint32_t)__popc(qmask) * BOX, kbytes = (uint32_t)__popc(rowmask) * BOX;
    auto issue_q = [&](int h) {
      mbar_arrive_expect_tx(qfull, qbytes);
#pragma unroll
      for (int i = 0; i < QR; ++i)
        if ((qmask >> i) & 1) tma_load_3d(sm_u + OFF_Q + i * BOX, &qmap, qfull, h * DH, tok_q[i], b);
    };
    int h = 0, ai = 0, ph = 0, st = 0, use = 0;
    long long prod_wait = 0;
    auto produce = [&](int s_end) {
      for (int s = (use * NST + st); s < s_end; ++s) {
        if (ph == 0 && ai == 0 && h >= 1) {  // every query-row warp holds the previous head's Q fragments in registers
          mbar_wait(qempty, (h - 1) & 1);
          issue_q(h);
        }
        if (use >= 1) {
          if (DBG && p.dbg != nullptr && blockIdx.x == 0) {
            const long long t0 = clock64();
            mbar_wait(&empty[st], (use - 1) & 1);
            prod_wait += clock64() - t0;
          } else {
            mbar_wait(&empty[st], (use - 1) & 1);
          }
        }
        const int ff = f - (kt - 1 - (a_lo + ai)) * p.dt;
        mbar_arrive_expect_tx(&full[st], kbytes);
        const int chan = (ph ? p.voff : p.koff) + h * DH;
        const int tokf = ff * GW * GW;
#pragma unroll
        for (int rr = 0; rr < NROW; ++rr)
          if ((rowmask >> rr) & 1) tma_load_3d(sm_u + st * STAGE + rr * BOX, &qmap, &full[st], chan, tokf + tok_k[rr], b);
        if (++st == NST) { st = 0; ++use; }
        if (++ai == n_a) { ai = 0; if (++h == NH) { h = 0; ph = 1; } }
      }
    };
    if (is_producer && lane == 0) {
      for (int i = 0; i < NST; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], QR);
      }
      mbar_init(qfull, 1);
      mbar_init(qempty, QR);
      fence_barrier_init();
      tma_prefetch_desc(&qmap);
      issue_q(0);
      produce(min(NST, n_a));  // head 0 only: later heads wait for the query-row warps, which start after the sync
    }
    if (!is_producer) one_time_state();
    cta_bar();  // barriers initialised (producer), shared state written (query-row warps)
    if (is_producer) {
      if (lane == 0) {
        produce(NS);
        if (DBG && p.dbg != nullptr && blockIdx.x == 0) p.dbg[23] = prod_wait;
      }
      return;
    }
  }


  // =============================== query-row warps ===============================
  const int yq = y0 + warp * p.dh;
  const int vbase = (f * GW + yq) * GW;                 // video index of this warp's x = 0
  const bool wactive = (yq < GW) && (vbase < p.nv);     // warp has at least one real query
  const bool ok0 = wactive && (vbase + g < p.nv), ok1 = wactive && (vbase + g + 8 < p.nv);
  bool blk_ok[KH];
#pragma unroll
  for (int bq = 0; bq < KH; ++bq) {
    const int yy = yq - (KH - 1 - bq) * p.dh;
    blk_ok[bq] = wactive && yy >= 0;
  }
  const bool all_ok = blk_ok[0];  // the earliest key row is in the grid -> every block is
  // the bos query (sequence row 0) attends only to itself (nuwa_pytorch.py:608): copy its value row
  if (f == 0 && y0 == 0 && warp == 0) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.o + (long long)b * p.o_bs);
    for (int c = lane; c < INNER / 2; c += 32) dst[c] = reinterpret_cast<const uint32_t*>(vbos)[c];
  }

  float* Sw = S32 + warp * GW * SP;
  __half* Pw = P16 + (size_t)warp * GW * PP;  // + h * QR*GW*PP
  // ---- per-lane band table: the 8 (query x, key x') pairs of a 16x16 block this lane holds, both as C fragment
  //      (scores) and as A fragment (probabilities): idx = nt*4 + e <-> x = g + 8*(e>>1), x' = nt*8 + 2t + (e&1).
  //      In-band entries address slot sbase + c; the others a trash slot (scores) / the always-zero slot (P). ----
  uint32_t s_addr[8], p_addr[8], inc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int x = g + 8 * ((i & 3) >> 1);
    const int xp = (i >> 2) * 8 + 2 * t + (i & 1);
    const int delta = x - xp;  // causal: key column = x - (kw-1-c)*dw
    int c = -1;
    if (delta >= 0 && delta % p.dw == 0 && delta / p.dw <= kw - 1) c = kw - 1 - delta / p.dw;
    inc[i] = c >= 0 ? 1u : 0u;
    // out-of-band entries go to a word private to this lane (a shared trash slot is a benign write-write race, but it
    // drowns compute-sanitizer racecheck in reports)
    s_addr[i] = c >= 0 ? smem_u32(Sw) + 4u * (x * SP + c) : sm_u + OFF_TRASH + 4u * (warp * 32 + lane);
    p_addr[i] = smem_u32(Pw) + 2u * (x * PP + (c >= 0 ? c : ZSLOT));
    // keep the tables in registers (ptxas otherwise rematerialises them in every block: +15 integer ops per block)
    asm volatile("" : "+r"(inc[i]), "+r"(s_addr[i]), "+r"(p_addr[i]));
  }
  // ldmatrix lane addressing inside a [16 rows x 128 B] SWIZZLE_128B box
  const int mat = lane >> 3, l7 = lane & 7;
  const int k_row = ((mat >> 1) << 3) + l7, k_ch = mat & 1;   // K (B operand): m0,m1 = keys 0-7 (k lo, hi); m2,m3 = keys 8-15
  const int a_row = ((mat & 1) << 3) + l7, a_ch = mat >> 1;   // Q (A operand) and V^T: m0,m1 = rows 0-7 / 8-15 of chunk c
  uint32_t k_sw[4], a_sw[4];                                  // swizzled byte offsets of the four 32-B chunk pairs
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    k_sw[i] = k_row * 128 + (((2 * i + k_ch) ^ l7) << 4);
    a_sw[i] = a_row * 128 + (((2 * i + a_ch) ^ l7) << 4);
    asm volatile("" : "+r"(k_sw[i]), "+r"(a_sw[i]));
  }
  const uint32_t stage0 = sm_u + warp * BOX;  // key row rr = warp + bq of a stage
  const int sbase_lo = 1 + a_lo * KH * kw, sbase_step = KH * kw;

  const bool dbg = DBG && p.dbg != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0;
  long long wait_cyc[2] = {0, 0};
  if (dbg) { p.dbg[0] = t_start; p.dbg[1] = clock64(); }
  auto wait_full = [&](int st, int par, int ph) {
    if (DBG && dbg) {
      const long long t0 = clock64();
      mbar_wait(&full[st], par);
      wait_cyc[ph] += clock64() - t0;
    } else {
      mbar_wait(&full[st], par);
    }
  };

  int st = 0, par = 0;
  // ================= phase 1: scores + softmax, head by head =================
  for (int h = 0; h < NH; ++h) {
    uint32_t qa[4][4];
    mbar_wait(qfull, h & 1);
    {
      const uint32_t qb = sm_u + OFF_Q + warp * BOX;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm4(qa[ks], qb + a_sw[ks]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(qempty);
    {  // bos key (slot 0): a block whose only non-zero key column is 0
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0 = 0u, b1 = 0u;
        if (g == 0) {
          b0 = *reinterpret_cast<const uint32_t*>(kbos + h * DH + ks * 16 + 2 * t);
          b1 = *reinterpret_cast<const uint32_t*>(kbos + h * DH + ks * 16 + 8 + 2 * t);
        }
        mma16816(c, qa[ks], b0, b1);
      }
      if (t == 0) {  // key column 0: c[0] (row g), c[2] (row g+8)
        Sw[g * SP] = c[0];
        Sw[(g + 8) * SP] = c[2];
      }
    }
    int sbase0 = sbase_lo;
    for (int ai = 0; ai < n_a; ++ai, sbase0 += sbase_step) {
      const uint32_t stage = stage0 + st * STAGE;
      wait_full(st, par, 0);
      if (all_ok) {
        // Hand-scheduled (the ldmatrix / mma / st.shared wrappers are volatile asm, so program order is issue
        // order): all operand loads of the KH blocks first, then the MMAs interleaved so that dependent ones are
        // KH issue slots apart, then the stores.  A warp issues in order; block-after-block costs a full
        // ldmatrix + 4-deep HMMA chain latency per block.
        uint32_t kf[KH][4][4];
#pragma unroll
        for (int bq = 0; bq < KH; ++bq)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ldsm4(kf[bq][ks], stage + bq * BOX + k_sw[ks]);
        float c0[KH][4], c1[KH][4];
#pragma unroll
        for (int bq = 0; bq < KH; ++bq)
#pragma unroll
          for (int e = 0; e < 4; ++e) c0[bq][e] = c1[bq][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int bq = 0; bq < KH; ++bq) {
            mma16816(c0[bq], qa[ks], kf[bq][ks][0], kf[bq][ks][1]);
            mma16816(c1[bq], qa[ks], kf[bq][ks][2], kf[bq][ks][3]);
          }
#pragma unroll
        for (int bq = 0; bq < KH; ++bq) {
          const uint32_t sb4 = 4u * (sbase0 + bq * kw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sts_f32(s_addr[e] + inc[e] * sb4, c0[bq][e]);
            sts_f32(s_addr[4 + e] + inc[4 + e] * sb4, c1[bq][e]);
          }
        }
      } else {
#pragma unroll
        for (int bq = 0; bq < KH; ++bq) {
          if (!blk_ok[bq]) continue;
          const uint32_t kb = stage + bq * BOX;
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t r[4];
            ldsm4(r, kb + k_sw[ks]);
            mma16816(c0, qa[ks], r[0], r[1]);
            mma16816(c1, qa[ks], r[2], r[3]);
          }
          const uint32_t sb4 = 4u * (sbase0 + bq * kw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sts_f32(s_addr[e] + inc[e] * sb4, c0[e]);
            sts_f32(s_addr[4 + e] + inc[4 + e] * sb4, c1[e]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done with the stage
      if (++st == NST) { st = 0; par ^= 1; }
    }
    // ---- head end: fp32 softmax of the 16 score rows -> P[h] (fp16); two lanes per row, interleaved slots.
    //      Slots J..47 hold the initial -FLT_MAX (never written) -> probability 0. ----
    {
      const int x = lane >> 1, half = lane & 1;
      const float* row = Sw + x * SP + half;
      float v[MAXJ / 2];
#pragma unroll
      for (int i = 0; i < MAXJ / 2; ++i) v[i] = row[2 * i];
      float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
      for (int i = 0; i < MAXJ / 2; ++i) m4[i & 3] = fmaxf(m4[i & 3], v[i]);
      float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      const float mneg = -m * p.scale_log2e;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < MAXJ / 2; ++i) {
        v[i] = fast_exp2(fmaf(v[i], p.scale_log2e, mneg));  // masked slots: exp2(-huge) == 0
        l4[i & 3] += v[i];
      }
      float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      const float inv = 1.0f / l;
      __half* prow = Pw + (size_t)h * QR * GW * PP + x * PP + half;
#pragma unroll
      for (int i = 0; i < MAXJ / 2; ++i) prow[2 * i] = __float2half_rn(v[i] * inv);
      __syncwarp();
    }
    if (dbg) p.dbg[2 + h] = clock64();
  }

  // ================= talking heads (nuwa_pytorch.py:556-558): P'[g][x][j] = sum_h W[g][h] P[h][x][j] =================
  {
    float W[NH * NH];
#pragma unroll
    for (int i = 0; i < NH * NH / 4; ++i) {
      const float4 w4 = reinterpret_cast<const float4*>(Wsm)[i];
      W[4 * i] = w4.x; W[4 * i + 1] = w4.y; W[4 * i + 2] = w4.z; W[4 * i + 3] = w4.w;
    }
    const int npair = (p.J + 1) >> 1;  // slot pairs, in place (each thread reads its 8 inputs before writing)
    for (int xx = lane >> 3; xx < GW; xx += 4)
      for (int jp = lane & 7; jp < npair; jp += 8) {
        uint32_t* base = reinterpret_cast<uint32_t*>(Pw + xx * PP + 2 * jp);
        float2 pin[NH];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const uint32_t u = base[hh * (QR * GW * PP / 2)];
          pin[hh] = __half22float2(*reinterpret_cast<const __half2*>(&u));
        }
#pragma unroll
        for (int gh = 0; gh < NH; ++gh) {
          float ax = 0.f, ay = 0.f;
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
            ax = fmaf(W[gh * NH + hh], pin[hh].x, ax);
            ay = fmaf(W[gh * NH + hh], pin[hh].y, ay);
          }
          base[gh * (QR * GW * PP / 2)] = pack_bf16x2(ax, ay);
        }
      }
    __syncwarp();
    if (dbg) p.dbg[10] = clock64();
  }

  // ================= phase 3: O[g] = P'[g] V[g], head by head =================
  for (int h = 0; h < NH; ++h) {
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    const uint32_t hoff = (uint32_t)h * (QR * GW * PP * 2);
    int sbase0 = sbase_lo;
    for (int ai = 0; ai < n_a; ++ai, sbase0 += sbase_step) {
      const uint32_t stage = stage0 + st * STAGE;
      wait_full(st, par, 1);
      if (all_ok) {  // hand-scheduled like phase 1: loads, then MMAs with 8 slots between dependent ones
        uint32_t av[KH][8];
#pragma unroll
        for (int bq = 0; bq < KH; ++bq) {
          const uint32_t sb2 = 2u * (sbase0 + bq * kw);
#pragma unroll
          for (int i = 0; i < 8; ++i) av[bq][i] = lds_u16(p_addr[i] + hoff + inc[i] * sb2);
        }
        uint32_t vf[KH][4][4];
#pragma unroll
        for (int bq = 0; bq < KH; ++bq)
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) ldsm4t(vf[bq][pr], stage + bq * BOX + a_sw[pr]);
#pragma unroll
        for (int bq = 0; bq < KH; ++bq) {
          uint32_t af[4];
          af[0] = av[bq][0] | (av[bq][1] << 16);  // row g,   keys 2t, 2t+1
          af[1] = av[bq][2] | (av[bq][3] << 16);  // row g+8, keys 2t, 2t+1
          af[2] = av[bq][4] | (av[bq][5] << 16);  // row g,   keys 2t+8, 2t+9
          af[3] = av[bq][6] | (av[bq][7] << 16);  // row g+8, keys 2t+8, 2t+9
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {
            mma16816(o[2 * pr], af, vf[bq][pr][0], vf[bq][pr][1]);
            mma16816(o[2 * pr + 1], af, vf[bq][pr][2], vf[bq][pr][3]);
          }
        }
      } else {
#pragma unroll
        for (int bq = 0; bq < KH; ++bq) {
          if (!blk_ok[bq]) continue;
          const uint32_t vb = stage + bq * BOX;
          const uint32_t sb2 = 2u * (sbase0 + bq * kw);
          uint32_t av[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) av[i] = lds_u16(p_addr[i] + hoff + inc[i] * sb2);
          uint32_t af[4];
          af[0] = av[0] | (av[1] << 16);
          af[1] = av[2] | (av[3] << 16);
          af[2] = av[4] | (av[5] << 16);
          af[3] = av[6] | (av[7] << 16);
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {
            uint32_t r[4];
            ldsm4t(r, vb + a_sw[pr]);
            mma16816(o[2 * pr], af, r[0], r[1]);
            mma16816(o[2 * pr + 1], af, r[2], r[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == NST) { st = 0; par ^= 1; }
    }
    if (wactive) {
      // ---- output head h: add the bos value (slot 0) and store ----
      const bf16* Pg = reinterpret_cast<const bf16*>(Pw + (size_t)h * QR * GW * PP);
      const float pb0 = __bfloat162float(Pg[g * PP]), pb1 = __bfloat162float(Pg[(g + 8) * PP]);
      bf16* ob = p.o + (long long)b * p.o_bs + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        const float2 vb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vbos + h * DH + nd * 8 + 2 * t));
        if (ok0)
          *reinterpret_cast<uint32_t*>(ob + (long long)(1 + vbase + g) * p.o_rs + nd * 8) =
              pack_bf16x2(fmaf(pb0, vb.x, o[nd][0]), fmaf(pb0, vb.y, o[nd][1]));
        if (ok1)
          *reinterpret_cast<uint32_t*>(ob + (long long)(1 + vbase + g + 8) * p.o_rs + nd * 8) =
              pack_bf16x2(fmaf(pb1, vb.x, o[nd][2]), fmaf(pb1, vb.y, o[nd][3]));
      }
    }
    if (dbg) p.dbg[11 + h] = clock64();
  }
  if (dbg) { p.dbg[19] = clock64(); p.dbg[20] = wait_cyc[0]; p.dbg[21] = wait_cyc[1]; p.dbg[22] = NS; }
}


long long* g_halo_dbg = nullptr;

}  // namespace

// measurement aid (tools/halo_stamps.py; not part of include/nuwa_b200.h): 24 int64 of device memory or NULL
extern "C" void nuwa_debug_halo_stamps(long long* dev_ptr) { g_halo_dbg = dev_ptr; }

// Envelope: causal full pass (t0 == 0, nq == nv + 1) over a 16-wide token grid, H == 8, dh == 64, kh <= 3,
// window <= 47 keys, q|k|v rows sharing one token stride.  NUWA_ERR_INVALID outside it (caller falls back).
int attn_3dna_halo(const AttnParams& p, cudaStream_t stream) {
  if (!p.causal || p.fmap != GW || p.t0 != 0 || p.t0_ptr != nullptr) return NUWA_ERR_INVALID;
  if (p.H != NH || p.dh != DH || p.nq != p.nv + 1 || p.nv <= 0 || p.B <= 0) return NUWA_ERR_INVALID;
  if (p.kh > 3 || p.kt < 1 || p.kh < 1 || p.kw < 1 || p.kw > GW || 1 + p.kt * p.kh * p.kw > MAXJ) return NUWA_ERR_INVALID;
  if (p.dt <= 0 || p.dh_ <= 0 || p.dw <= 0) return NUWA_ERR_INVALID;
  // causal padding must equal dil*(k-1) exactly (nuwa_pytorch.py:424-428 pads 2*(dil*(k-1)//2))
  if (((p.dt * (p.kt - 1)) & 1) || ((p.dh_ * (p.kh - 1)) & 1) || ((p.dw * (p.kw - 1)) & 1)) return NUWA_ERR_INVALID;
  if (p.head_scale != nullptr || p.bias != nullptr || p.key_mask != nullptr || p.null_k != nullptr) return NUWA_ERR_INVALID;
  const bf16* q = reinterpret_cast<const bf16*>(p.q);
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* v = reinterpret_cast<const bf16*>(p.v);
  const long long koff = k - q, voff = v - q;
  if (p.k_rs != p.q_rs || p.v_rs != p.q_rs || p.k_bs != p.q_bs || p.v_bs != p.q_bs) return NUWA_ERR_INVALID;
  if (koff < 0 || voff < 0 || koff + INNER > p.q_rs || voff + INNER > p.q_rs) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.q_bs % 8) || (koff % 8) || (voff % 8) || (p.o_rs & 1) || (p.o_bs & 1)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.o) & 3)) return NUWA_ERR_INVALID;

  CUtensorMap map;
  const uint64_t dims[3] = {(uint64_t)p.q_rs, (uint64_t)(p.nv + 1), (uint64_t)p.B};
  const uint64_t strides[3] = {2, (uint64_t)p.q_rs * 2, (uint64_t)p.q_bs * 2};
  const uint32_t box[3] = {DH, GW, 1};
  const int rc = encode_map_bf16_sw128(&map, p.q, 3, dims, strides, box);
  if (rc != NUWA_OK) return rc;

  HaloArgs a;
  a.B = p.B; a.nv = p.nv;
  a.nf = (p.nv + GW * GW - 1) / (GW * GW);
  int tiles = 0;
  for (int r = 0; r < p.dh_ && r < GW; ++r) tiles += ((GW - r + p.dh_ - 1) / p.dh_ + QR - 1) / QR;
  a.tiles = tiles;
  a.kt = p.kt; a.kh = p.kh; a.kw = p.kw; a.dt = p.dt; a.dh = p.dh_; a.dw = p.dw;
  a.koff = (int)koff; a.voff = (int)voff;
  a.J = 1 + p.kt * p.kh * p.kw;
  a.scale_log2e = p.qscale * 1.4426950408889634f;
  a.talk = p.talk;
  a.o = reinterpret_cast<bf16*>(p.o); a.o_bs = p.o_bs; a.o_rs = p.o_rs;
  a.k0 = k; a.v0 = v; a.k_bs = p.k_bs; a.v_bs = p.v_bs;
  a.dbg = g_halo_dbg;

  const int grid = a.nf * a.tiles * a.B;
  auto launch = [&](void (*kern)(const CUtensorMap, const HaloArgs), int slot) -> int {
    (void)slot;  // (attributes are set per launch: idempotent driver calls, no cached library state)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return NUWA_ERR_CUDA;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<grid, (QR + 1) * 32, SMEM_BYTES, stream>>>(map, a);
    return NUWA_OK;
  };
  int lrc;
  if (a.dbg != nullptr && p.kh == 3) lrc = launch(attn_3dna_halo_kernel<3, true>, 4);  // tools/halo_stamps.py
  else if (p.kh == 1) lrc = launch(attn_3dna_halo_kernel<1, false>, 1);
  else if (p.kh == 2) lrc = launch(attn_3dna_halo_kernel<2, false>, 2);
  else lrc = launch(attn_3dna_halo_kernel<3, false>, 3);
  if (lrc != NUWA_OK) return lrc;
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace nuwa
