// Sparse3DNA attention core on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, K / V key rows
// staged by TMA, one CTA per SM walking a cost-sorted tile list.
//
// Reference: Sparse3DNA.forward core, nuwa_pytorch.py:523-564 (unfoldNd gather of k,v -> einsum -> mask -> softmax ->
// talking heads -> einsum), causal (decoder) and centred (NUWASketch sketch encoder) windows.
//
// Work decomposition.  A TILE is 128 queries = 8 token-grid rows x 16 columns of one frame of one sample, all 8 heads.
// The 8 rows are spaced by the window's row dilation (rows r, r+dh, r+2dh, ...; for dh >= 4 two such classes of 4 rows),
// so the key rows the tile needs per frame offset are only 8 + kh - 1 <= 10 "slot columns" of 16 keys: the B operand of
// ONE UMMA (N = 16 x valid slot columns <= 160), fetched by one or two TMA boxes over a 5-D view of the q|k|v buffer whose
// row axis is split by the dilation.  A UNIT is (head, frame offset a):
//   phase 1   S = Q_h K_{h,a}^T   M=128, N<=160, K=64, accumulator in TMEM (one S buffer per warpgroup); thread = query
//             row: tcgen05.ld of the 64-column window holding its 3 key rows, dumped to a thread-private shared-memory
//             row (16 vector stores), the 9 in-band entries read back at per-tile precomputed offsets into registers; the
//             bos score is a 64-term dot product per thread straight from the Q tile.  Per head: fp32 softmax ->
//             fp16 probabilities parked in TMEM, column 8u + h = slots (2u, 2u+1) of head h
//   mix       talking heads (nuwa_pytorch.py:556-558) ON THE TENSOR CORES: for every slot pair u one UMMA whose A operand
//             is read from tensor memory (the 8 heads x 2 slots = K 16 just parked there) against the 16 x 16 block
//             matrix [W 0; 0 W] split into fp16 high + low parts (two accumulating UMMAs: W to ~22 bits); the fp32 result
//             is rounded to bf16 P' and written back in place
//   phase 2   P'_g placed at its key position inside a dense 128 x 160 A operand that lives in TENSOR MEMORY
//             (tcgen05.st; byte-permute placement, no shared-memory traffic), O_g += P'_g V_{g,a}: A from TMEM, B = the
//             V key rows as MN-major operand, N=64; per head: + bos probability x bos value -> bf16 -> global
// Warp roles: warps 0 / 11 = TMA producers (one per warpgroup), warps 2-5 / 6-9 = two warpgroups that own the even / odd heads, so extraction +
// softmax of one head overlaps the UMMAs of the other; warp 1 / warp 10 = one MMA-issuing thread per warpgroup (a
// tcgen05.mma costs its issuing thread ~70 cycles, the small UMMAs of this kernel are issue bound from one thread).
// TMEM columns: P [0,184);  phase 1  S0 [192,352) S1 [352,512);  mix  D [192,512);  phase 2  A0 [192,272) A1 [272,352)
// O0 [352,416) O1 [416,480).
//
// Variants (compile time): X2 = SparseCross2DNA (nuwa_pytorch.py:851-895: queries and context in separate buffers, unit =
// context frame, learned fp32 null key / value in slot 0, context mask on the gathered scores).  MODE 1 = scores only
// (phase 1, band written to HBM as fp32: logits and dP' = dO V^T of the backward pass), MODE 2 = PV only (phase 2 with a
// bf16 tensor from HBM in the place of P' and V := K: dq of the backward pass).
#include <float.h>
#include <cuda_fp16.h>

#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "tmem_ldst.cuh"

namespace nuwa {

int encode_map_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);  // gemm_tcgen05.cu

namespace {

constexpr int GW = 16;                 // token grid width == height
constexpr int NH = 8, DH = 64, INNER = NH * DH;
constexpr int KW = 3;                  // window width (envelope: kw == 3)
constexpr int MAXKT = 5, MAXKH = 3;
constexpr int MAXJ = 1 + MAXKT * MAXKH * KW;   // 46
constexpr int NU = (MAXJ - 1 + 1) / 2;         // 23 slot pairs
constexpr int NSC = 8 + MAXKH - 1;     // slot columns of a key tile (10)
constexpr int BOX = GW * DH * 2;       // 2048 B: 16 tokens x 64 channels of one head
constexpr int KV_STAGE = NSC * BOX;    // 20 KB
constexpr int NST = 4;
constexpr int Q_BYTES = 8 * BOX;       // 16 KB
constexpr int DPITCH = 272;            // bytes per thread row of the window dump: 64 fp32 + a -FLT_MAX sentinel, 16-B aligned

constexpr int OFF_KV = 0;
constexpr int OFF_Q = OFF_KV + NST * KV_STAGE;            // 4 buffers: (head pair parity, warpgroup)
constexpr int OFF_KBOS = OFF_Q + 4 * Q_BYTES;             // [8 heads][64] fp32: bos key rows of the tile's sample
constexpr int OFF_WB = OFF_KBOS + NH * DH * 4;            // talking-heads B operands: [hi, lo] x [16 rows x 128 B]
constexpr int OFF_SC = OFF_WB + 2 * BOX;                  // 2 warpgroups x [128] window dump rows (thread private)
constexpr int OFF_VBOS = OFF_SC + 2 * 128 * DPITCH;       // [8][64] fp32: value of slot 0 (bos row / learned null value)
constexpr int OFF_PBOS = OFF_VBOS + INNER * 4;            // [8][128] fp32: probability of slot 0
constexpr int OFF_W = OFF_PBOS + NH * 128 * 4;            // [8][8] fp32
constexpr int OFF_BAR = OFF_W + NH * NH * 4;
constexpr int NBAR = 2 * NST + 2 * 4 + 6 * 2 + 3;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;

constexpr int T_P = 0, T_S = 192, T_D = 192, T_A = 192, T_O = 352;
constexpr int THREADS = 384;   // warps 0 / 11: TMA producers, warps 1 / 10: MMA issuers of warpgroup 0 / 1, warps 2-9: warpgroups

struct UmmaArgs {
  int B, nv, nf, maxf, tpf, ntiles;    // nf = frames present, tpf = tiles per frame
  int kt, kh, dt, dh, dw, causal;
  int koff, voff;                      // channel offsets of k / v inside a row of the key buffer
  int q_rows5d, kv_rows5d;             // rows [0, rows5d) of every sample are reachable through the 5-D maps (0: none)
  int q_tok0, kv_tok0;                 // token index of grid position 0 in the flat query / key views (1: a bos row first)
  int abs_frames;                      // SparseCross2DNA: frame offset a addresses context frame a (else f + (a - At) dt)
  const float* null_k;                 // learned null key / value [H * dh] fp32 (SparseCross2DNA) or NULL: key row 0 (bos)
  const float* null_v;
  const unsigned char* key_mask;       // [B][mask_bs] context-token mask or NULL
  int mask_bs;
  float scale_log2e;
  const float* talk;
  bf16* o;
  long long o_bs;
  int o_rs;
  const bf16* k0;
  const bf16* v0;
  long long k_bs, v_bs;
  // scores mode (backward): phase 1 only, the band scores go to s_out[b][h][query][s_jp] as fp32 (slot j = 1 + (a kh + b) 3 + c,
  // slot 0 = bos / null key), multiplied by s_scale; masked slots get s_masked
  float* s_out;
  const bf16* pv_in;   // PV mode: dS [B][H][nq][s_jp] (slot order, slot 0 = bos / null key)
  int s_jp;
  float s_scale, s_masked;
  long long* dbg;  // tools/umma_stamps.py: clock64 stamps of CTA 0 (NULL in normal use)
};

// geometry of one tile (registers only), identical in every role
struct Tile {
  int b, f;
  int two;           // two classes of 4 rows (row dilation >= 4)
  int r0, i0;        // single class: rows r0 + (i0 + i) * dh;  two classes: classes r0 and r0 + 1
  int sc_lo, sc_hi;  // valid slot columns [sc_lo, sc_hi)
  int real_mask;     // frame offsets a with a real key frame
  int zero_mask;     // frame offsets whose frame lies inside the volume but beyond the sequence (visible zero keys, D16)
  int n_real;
};

// grid row of tile row i (-1: none)
__device__ __forceinline__ int tile_row(const UmmaArgs& p, const Tile& t, int i) {
  if (!t.two) {
    const int y = t.r0 + (t.i0 + i) * p.dh;
    return y < GW ? y : -1;
  }
  const int r = t.r0 + (i >> 2);
  const int y = r + (i & 3) * p.dh;
  return (r < p.dh && y < GW) ? y : -1;
}
// grid row of slot column sc (-1: none); Ah = rows of the window above the query row
__device__ __forceinline__ int slot_row(const UmmaArgs& p, const Tile& t, int sc, int Ah) {
  if (!t.two) {
    const int y = t.r0 + (t.i0 + sc - Ah) * p.dh;
    return (y >= 0 && y < GW && sc < 8 + p.kh - 1) ? y : -1;
  }
  const int sl = sc - Ah;
  if (sl < 0 || sl >= 8) return -1;
  const int r = t.r0 + (sl >> 2);
  const int y = r + (sl & 3) * p.dh;
  return (r < p.dh && y < GW) ? y : -1;
}

__device__ __forceinline__ void make_tile(const UmmaArgs& p, int s, int Ah, Tile& t) {
  // sorted tile index s: later frames first (causal: more key frames = more work)
  const int per_f = p.tpf * p.B;
  t.f = p.nf - 1 - s / per_f;
  const int rem = s % per_f;
  const int ti = rem / p.B;
  t.b = rem - ti * p.B;
  t.two = p.dh >= 4;
  if (!t.two) {
    int r = 0, k = ti;
    for (; r < p.dh; ++r) {
      const int nrows = (GW - r + p.dh - 1) / p.dh;
      const int nt = (nrows + 7) / 8;
      if (k < nt) break;
      k -= nt;
    }
    t.r0 = r;
    t.i0 = k * 8;
  } else {
    t.r0 = 2 * ti;
    t.i0 = 0;
  }
  t.sc_lo = NSC; t.sc_hi = 0;
#pragma unroll
  for (int sc = 0; sc < NSC; ++sc)
    if (slot_row(p, t, sc, Ah) >= 0) { t.sc_lo = min(t.sc_lo, sc); t.sc_hi = max(t.sc_hi, sc + 1); }
  const int At = p.causal ? p.kt - 1 : (p.kt - 1) / 2;
  t.real_mask = t.zero_mask = t.n_real = 0;
#pragma unroll
  for (int a = 0; a < MAXKT; ++a) {
    if (p.abs_frames) {                  // every context frame is a real key frame for every query frame
      if (a < p.kt) { t.real_mask |= 1 << a; ++t.n_real; }
      continue;
    }
    const int ff = t.f + (a - At) * p.dt;
    if (a < p.kt && ff >= 0 && ff < p.maxf) {
      if (ff < p.nf) { t.real_mask |= 1 << a; ++t.n_real; }
      else t.zero_mask |= 1 << a;
    }
  }
}

// k-th tile of CTA c under the snake deal of the cost-sorted list: rounds alternate direction
__device__ __forceinline__ int snake(int k, int c, int G) { return k * G + ((k & 1) ? G - 1 - c : c); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// key / query rows y0 + i * dh (i < n) of frame f, channels [chan, chan + 64) -> dst + i * BOX.  Rows reachable through
// the 5-D view travel as boxes of 8 / 4 / 2 rows (one TMA instruction each); the ragged end of the sequence and
// dilations that do not divide the grid height use one 16-token box per row on the flat view (zero fill past the end).
struct MapSet {
  const CUtensorMap* flat;
  const CUtensorMap* m8;
  const CUtensorMap* m4;
  const CUtensorMap* m2;
  int rows5d, tok0;
};
__device__ __forceinline__ void load_rows(uint32_t dst, uint64_t* bar, const MapSet& ms, int dh, int chan, int f, int y0, int n,
                                          int b) {
  const int R0 = f * GW + y0;
  int i = 0;
  if (R0 + (n - 1) * dh < ms.rows5d) {
    const int ylo = y0 % dh, yhi = R0 / dh;
    for (; n - i >= 8; i += 8) tma_load_5d(dst + i * BOX, ms.m8, bar, chan, 0, ylo, yhi + i, b);
    if (n - i >= 4) { tma_load_5d(dst + i * BOX, ms.m4, bar, chan, 0, ylo, yhi + i, b); i += 4; }
    if (n - i >= 2) { tma_load_5d(dst + i * BOX, ms.m2, bar, chan, 0, ylo, yhi + i, b); i += 2; }
  }
  for (; i < n; ++i) tma_load_3d(dst + i * BOX, ms.flat, bar, chan, ms.tok0 + (R0 + i * dh) * GW, b);
}

// X2 = SparseCross2DNA variant (absolute context frames, learned null key / value, context mask): a compile-time switch so
// that the Sparse3DNA instantiation carries none of its per-unit work (mask selects, 64-bit mask word) or registers.
// SC = scores mode: S = Q K^T band extraction only (no softmax / mix / PV); used twice by the backward pass, for the logits
// (Q = q, K = k) and for dP' = dO V^T (Q = dO from its own buffer, K = v).
// MODE: 0 = forward, 1 = scores (see SC above), 2 = PV only: the probabilities' place is taken by a bf16 tensor read from HBM
// (dS of the backward pass, [B][H][nq][jp] in slot order) and V := K, so the output is dq = sum_j dS[j] k_j (slot 0 uses
// the bos / null key); no Q tiles, no phase 1, no softmax, no mix.
template <int DW, bool X2, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
attn_3dna_umma_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap qmap8,
                      const __grid_constant__ CUtensorMap qmap4, const __grid_constant__ CUtensorMap qmap2,
                      const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap kmap8,
                      const __grid_constant__ CUtensorMap kmap4, const __grid_constant__ CUtensorMap kmap2, const UmmaArgs p) {
  constexpr bool SC = MODE == 1, PV = MODE == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* full = bars;                 // [NST]  TMA -> MMA
  uint64_t* empty = full + NST;          // [NST]  MMA commit -> TMA
  uint64_t* qfull = empty + NST;         // [4]    Q tile of (head pair parity, warpgroup)
  uint64_t* qempty = qfull + 4;          // [4]
  uint64_t* sfull = qempty + 4;          // [2]    MMA commit -> warpgroup
  uint64_t* sempty = sfull + 2;          // [2]    warpgroup (4 warps) -> MMA
  uint64_t* afull = sempty + 2;          // [2]    warpgroup -> MMA (A operand written to TMEM)
  uint64_t* aempty = afull + 2;          // [2]    MMA commit -> warpgroup
  uint64_t* ofull = aempty + 2;          // [2]
  uint64_t* oempty = ofull + 2;          // [2]
  uint64_t* tready = oempty + 2;         // [1]    8 warps: previous tile drained, this tile's bos rows staged
  uint64_t* mixgo = tready + 1;          // [1]    8 warps -> MMA: probabilities parked / first batch converted
  uint64_t* mixfull = mixgo + 1;         // [1]    MMA commit -> warpgroups: a batch of mixed probabilities is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&qfull[i], 1); mbar_init(&qempty[i], 5); }  // commit + 4 warps (bos dot)
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 4);
      mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1);
      mbar_init(&ofull[i], 1); mbar_init(&oempty[i], 4);
    }
    mbar_init(tready, 8);
    mbar_init(mixgo, 8);
    mbar_init(mixfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&qmap);
    tma_prefetch_desc(&qmap8);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&kmap8);
    tma_prefetch_desc(&kmap2);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // one-time shared-memory state: operand buffers zeroed (slot columns that are never loaded must hold finite values:
  // their probabilities are exactly 0, and 0 x NaN would poison P'V), bos operand tiles, talking-heads operands
  for (int i = tid; i < OFF_SC / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < NH * NH) {
    float* Wsm = reinterpret_cast<float*>(sm + OFF_W);
    const float w = p.talk ? __ldg(p.talk + tid) : ((tid / NH) == (tid % NH) ? 1.f : 0.f);
    Wsm[tid] = w;
    // B[n = 2g + s'][k = 2h + s] = W[g][h] (s == s'), fp16 high and low parts; rows of 128 B, SWIZZLE_128B chunks
    const int g = tid / NH, h = tid % NH;
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
#pragma unroll
    for (int sp = 0; sp < 2; ++sp) {
      const int n = 2 * g + sp, kk = 2 * h + sp;
      const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 7u) * 2u;
      *reinterpret_cast<__half*>(sm + OFF_WB + off) = hi;
      *reinterpret_cast<__half*>(sm + OFF_WB + BOX + off) = lo;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int Ah = p.causal ? p.kh - 1 : (p.kh - 1) / 2;
  const int At = p.causal ? p.kt - 1 : (p.kt - 1) / 2;

  if (warp == 0 || warp == 11) {
    // ============ TMA producer of warpgroup w: its Q tiles and every second ring slot (its K / V tiles) ============
    if (lane == 0) {
      const int w = warp == 0 ? 0 : 1;
      const MapSet qms = {&qmap, &qmap8, &qmap4, &qmap2, p.q_rows5d, p.q_tok0};
      const MapSet kms = {&kmap, &kmap8, &kmap4, &kmap2, p.kv_rows5d, p.kv_tok0};
      uint32_t it = (uint32_t)w, qu[2] = {0, 0};
      for (int k = 0;; ++k) {
        const int s = snake(k, cta, G);
        if (s >= p.ntiles) break;
        Tile t;
        make_tile(p, s, Ah, t);
        // row runs: queries (from tile row 0 / 4) and key slot columns (from sc_lo, or Ah / Ah + 4 for two classes)
        int qy[2] = {0, 0}, qn[2] = {0, 0}, ky[2] = {0, 0}, kn[2] = {0, 0}, ksc[2] = {0, 0};
        if (!t.two) {
          qy[0] = tile_row(p, t, 0);
          for (int i = 0; i < 8; ++i) qn[0] += tile_row(p, t, i) >= 0;
          ksc[0] = t.sc_lo;
          ky[0] = slot_row(p, t, t.sc_lo, Ah);
          kn[0] = t.sc_hi - t.sc_lo;
        } else {
          for (int c = 0; c < 2; ++c) {
            qy[c] = t.r0 + c;
            for (int i = 0; i < 4; ++i) qn[c] += tile_row(p, t, 4 * c + i) >= 0;
            ksc[c] = Ah + 4 * c;
            ky[c] = t.r0 + c;
            kn[c] = qn[c];
          }
        }
        const uint32_t qbytes = (uint32_t)(qn[0] + qn[1]) * BOX;
        const uint32_t kbytes = (uint32_t)(kn[0] + kn[1]) * BOX;
        // fast path of the key tiles (every row reachable through the 5-D view): run c = box of 8 rows + box of 2 / 4
        // rows, coordinates precomputed up to the frame term
        const int rpf = p.kv_rows5d > 0 ? GW / p.dh : 0;               // 5-D row-block coordinate advance per frame
        int k_ylo[2], k_yhi[2], k_last[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          k_ylo[c] = p.kv_rows5d > 0 ? ky[c] % p.dh : 0;
          k_yhi[c] = p.kv_rows5d > 0 ? ky[c] / p.dh : 0;
          k_last[c] = ky[c] + (kn[c] - 1) * p.dh;                       // last row inside its frame
        }
        auto load_q = [&](int hp) {
          const int h = 2 * hp + w, qb = 2 * (hp & 1) + w;
          mbar_wait(&qempty[qb], (qu[hp & 1] & 1) ^ 1);
          mbar_arrive_expect_tx(&qfull[qb], qbytes);
          for (int c = 0; c < 2; ++c)
            if (qn[c] > 0)
              load_rows(sm_u + OFF_Q + qb * Q_BYTES + 4 * c * BOX, &qfull[qb], qms, p.dh, h * DH, t.f, qy[c], qn[c], t.b);
          ++qu[hp & 1];
        };
        if constexpr (!PV) load_q(0);
        for (int ph = (PV ? 1 : 0); ph < (SC ? 1 : 2); ++ph) {
          for (int hp = 0; hp < NH / 2; ++hp) {
            const int chan = (ph ? p.voff : p.koff) + (2 * hp + w) * DH;
            for (int a = 0; a < p.kt; ++a) {
              if (!((t.real_mask >> a) & 1)) continue;
              const int ff = X2 ? a : t.f + (a - At) * p.dt;
              const int st = it % NST;
              const uint32_t dst = sm_u + OFF_KV + st * KV_STAGE;
              mbar_wait(&empty[st], ((it / NST) & 1) ^ 1);
              mbar_arrive_expect_tx(&full[st], kbytes);
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                if (kn[c] <= 0) continue;
                const uint32_t d = dst + ksc[c] * BOX;
                if (ff * GW + k_last[c] < p.kv_rows5d && (kn[c] == 10 || kn[c] == 8 || kn[c] == 4)) {
                  const int yhi = ff * rpf + k_yhi[c];
                  if (kn[c] == 4) {
                    tma_load_5d(d, &kmap4, &full[st], chan, 0, k_ylo[c], yhi, t.b);
                  } else {
                    tma_load_5d(d, &kmap8, &full[st], chan, 0, k_ylo[c], yhi, t.b);
                    if (kn[c] == 10) tma_load_5d(d + 8 * BOX, &kmap2, &full[st], chan, 0, k_ylo[c], yhi + 8, t.b);
                  }
                } else {
                  load_rows(d, &full[st], kms, p.dh, chan, ff, ky[c], kn[c], t.b);
                }
              }
              it += 2;
            }
            // the next head pair's Q tile goes out behind this pair's key tiles: its buffer was released a whole
            // head pair ago, and it lands long before the issuer gets there
            if (!PV && ph == 0 && hp + 1 < NH / 2) load_q(hp + 1);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ====================== MMA issuer of warpgroup w (w = 0 also issues the talking-heads UMMAs) ======================
    if (lane == 0) {
      const int w = warp == 1 ? 0 : 1;
      uint32_t it = (uint32_t)w, qu[2] = {0, 0}, su = 0, au = 0, ou = 0, tile_n = 0, mg = 0;
      const uint32_t idesc_pv = make_idesc_f16(128, DH, 1, 1, 0, 1);   // A: bf16 from TMEM, B: bf16 MN-major
      const uint32_t idesc_mix = make_idesc_f16(128, 16, 0, 0, 0, 0);  // A: fp16 from TMEM, B: fp16 K-major
      const uint64_t wb_hi = make_sw128_kmajor_desc(sm_u + OFF_WB), wb_lo = make_sw128_kmajor_desc(sm_u + OFF_WB + BOX);
      const uint32_t d_s = tmem + T_S + w * 160, d_o = tmem + T_O + w * DH, a_t = tmem + T_A + w * 80;
      for (int k = 0;; ++k) {
        const int s = snake(k, cta, G);
        if (s >= p.ntiles) break;
        Tile t;
        make_tile(p, s, Ah, t);
        const int nsc = t.sc_hi - t.sc_lo;
        const uint32_t idesc_qk = make_idesc_f16(128, 16 * nsc, 1, 1, 0, 0);
        mbar_wait(tready, tile_n & 1);
        ++tile_n;
        tc_fence_after();
        // ---------------- phase 1: S = Q K^T ----------------
        for (int hp = 0; hp < (PV ? 0 : NH / 2); ++hp) {
          const int qb = 2 * (hp & 1) + w;
          const uint64_t qdesc = make_sw128_kmajor_desc(sm_u + OFF_Q + qb * Q_BYTES);
          mbar_wait(&qfull[qb], qu[hp & 1] & 1);
          for (int a = 0; a < p.kt; ++a) {
            if (!((t.real_mask >> a) & 1)) continue;
            const int st = it % NST;
            const uint64_t kdesc = make_sw128_kmajor_desc(sm_u + OFF_KV + st * KV_STAGE + t.sc_lo * BOX);
            const uint32_t d = d_s + t.sc_lo * 16;
            const bool md = p.dbg != nullptr && cta == 0 && k == 0 && w == 0 && su < 24;
            if (md) p.dbg[320 + 4 * su] = clock64();
            mbar_wait2(&full[st], (it / NST) & 1, &sempty[w], (su & 1) ^ 1);
            if (md) { p.dbg[320 + 4 * su + 1] = clock64(); p.dbg[320 + 4 * su + 2] = p.dbg[320 + 4 * su + 1]; }
            tc_fence_after();
            umma_bf16(d, qdesc, kdesc, idesc_qk, 0u);
            umma_bf16(d, qdesc + 2, kdesc + 2, idesc_qk, 1u);
            umma_bf16(d, qdesc + 4, kdesc + 4, idesc_qk, 1u);
            umma_bf16(d, qdesc + 6, kdesc + 6, idesc_qk, 1u);
            umma_commit(&empty[st]);
            umma_commit(&sfull[w]);
            if (md) p.dbg[320 + 4 * su + 3] = clock64();
            it += 2; ++su;
          }
          umma_commit(&qempty[qb]);
          ++qu[hp & 1];
        }
        if constexpr (SC) continue;
        // ---------------- talking heads: D[q][2g + s'] = sum_h W[g][h] P[h][2u + s'] for every slot pair u ----------------
        if (!PV && w == 0) {
          for (int batch = 0; batch < 2; ++batch) {
            mbar_wait(mixgo, mg & 1);
            ++mg;
            tc_fence_after();
            const int u0 = batch ? 20 : 0, u1 = batch ? NU : 20;
            for (int u = u0; u < u1; ++u) {
              umma_f16_ts(tmem + T_D + 16 * (u - u0), tmem + T_P + 8 * u, wb_hi, idesc_mix, 0u);
              umma_f16_ts(tmem + T_D + 16 * (u - u0), tmem + T_P + 8 * u, wb_lo, idesc_mix, 1u);
            }
            umma_commit(mixfull);
          }
        }
        // ---------------- phase 2: O = P' V ----------------
        for (int gp = 0; gp < NH / 2; ++gp) {
          uint32_t acc = 0u;
          int seen = 0;
          for (int a = 0; a < p.kt; ++a) {
            if (!((t.real_mask >> a) & 1)) continue;
            ++seen;
            const int st = it % NST;
            const uint64_t vdesc = make_sw128_kmajor_desc(sm_u + OFF_KV + st * KV_STAGE);
            const bool md = p.dbg != nullptr && cta == 0 && k == 0 && w == 0 && au < 24;
            if (md) p.dbg[192 + 4 * au] = clock64();
            if (seen == 1) mbar_wait(&oempty[w], (ou & 1) ^ 1);
            mbar_wait2(&full[st], (it / NST) & 1, &afull[w], au & 1);
            if (md) { p.dbg[192 + 4 * au + 1] = clock64(); p.dbg[192 + 4 * au + 2] = p.dbg[192 + 4 * au + 1]; }
            tc_fence_after();
#pragma unroll
            for (int sc = 0; sc < NSC; ++sc)
              if (sc >= t.sc_lo && sc < t.sc_hi) {
                umma_f16_ts(d_o, a_t + sc * 8, vdesc + (uint64_t)(sc * (BOX >> 4)), idesc_pv, acc);
                acc = 1u;
              }
            umma_commit(&empty[st]);
            umma_commit(&aempty[w]);
            if (seen == t.n_real) { umma_commit(&ofull[w]); ++ou; }
            if (md) p.dbg[192 + 4 * au + 3] = clock64();
            it += 2; ++au;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== warpgroups ===========================================
    const int wg = (warp - 2) >> 2;              // 0: even heads, 1: odd heads
    const int quarter = warp & 3;                // TMEM lane quarter this warp may touch
    const int qrow = quarter * 32 + lane;        // query index inside the tile == TMEM lane
    const int ti = qrow >> 4, x = qrow & 15;     // tile row, grid column
    const int r = ti & 1;                        // row inside the warp's pair
    const int wq = quarter;                      // window index: slot columns [2*wq, 2*wq + 4)
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t drow_u = sm_u + OFF_SC + (uint32_t)(wg * 128 + qrow) * DPITCH;   // this thread's window dump row
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(drow_u + 256), "f"(-FLT_MAX) : "memory");  // sentinel for masked slots
    const float* Wsm = reinterpret_cast<const float*>(sm + OFF_W);
    const float* vbos = reinterpret_cast<const float*>(sm + OFF_VBOS);
    float* pbos_s = reinterpret_cast<float*>(sm + OFF_PBOS);
    const int Aw = p.causal ? KW - 1 : (KW - 1) / 2;

    // ---- per-thread constants of the key-column geometry (x only) ----
    bool vc[KW];
    uint32_t sel[8];                             // PRMT selectors placing (P'_c0, P'_c1 | P'_c2, 0) at their key columns
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sel[w8] = 0x7676u;
#pragma unroll
    for (int c = 0; c < KW; ++c) {
      const int jc = x + (c - Aw) * p.dw;
      vc[c] = jc >= 0 && jc < GW;
      if (vc[c]) {
        const uint32_t nib = (uint32_t)(2 * c) | ((uint32_t)(2 * c + 1) << 4);   // bytes of half-word c
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8)
          if ((jc >> 1) == w8) sel[w8] = (jc & 1) ? ((sel[w8] & 0x00ffu) | (nib << 8)) : ((sel[w8] & 0xff00u) | nib);
      }
    }
    uint32_t su = 0, au = 0, ou = 0, mf = 0, qu = 0;   // qu: uses of this warpgroup's Q buffers (same count for both parities)

    for (int k = 0;; ++k) {
      const int s = snake(k, cta, G);
      if (s >= p.ntiles) break;
      Tile t;
      make_tile(p, s, Ah, t);
      const int y = tile_row(p, t, ti);
      const int vpos = (t.f * GW + (y >= 0 ? y : 0)) * GW + x;      // video-token index of this query
      const bool qok = y >= 0 && vpos < p.nv;
      // key-row validity of this thread's kernel rows, window extraction mask
      uint32_t vb = 0;
#pragma unroll
      for (int b = 0; b < MAXKH; ++b) {
        const int sc = ti + b;
        bool ok = b < p.kh && slot_row(p, t, sc, Ah) >= 0;
        if (t.two) { const int sl = sc - Ah; ok = ok && sl >= 0 && (sl >> 2) == (ti >> 2); }
        vb |= (ok ? 1u : 0u) << b;
      }
      // byte offsets, inside the dump row, of the 9 window entries (kernel row b, kernel column c) of this thread: window
      // column 16 * (r + b) + x + (c - Aw) * dw; masked entries point at the -FLT_MAX sentinel.  The same for every unit.
      uint32_t goff[MAXKH * KW];
      uint32_t valid9 = 0;
#pragma unroll
      for (int b = 0; b < MAXKH; ++b)
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const bool ok = ((vb >> b) & 1u) && vc[c];
          goff[b * KW + c] = ok ? 4u * (uint32_t)(16 * (r + b) + x + (c - Aw) * p.dw) : 256u;
          valid9 |= (ok ? 1u : 0u) << (b * KW + c);
        }
      // context-token mask (SparseCross2DNA, nuwa_pytorch.py:878-884): the window of a query sits at the same grid position
      // in every context frame, so its 9 mask bits per frame are tile constants
      unsigned long long kmask = 0;      // 9 bits per frame offset (X2 only)
      if constexpr (X2) {
#pragma unroll
        for (int a = 0; a < MAXKT; ++a) {
          uint32_t m = valid9;
          if (p.key_mask != nullptr && a < p.kt) {
            m = 0;
#pragma unroll
            for (int b = 0; b < MAXKH; ++b)
#pragma unroll
              for (int c = 0; c < KW; ++c)
                if ((valid9 >> (b * KW + c)) & 1u) {
                  const int yy = slot_row(p, t, ti + b, Ah), xx = x + (c - Aw) * p.dw;   // valid9: inside the grid
                  const int tok = (a * GW + yy) * GW + xx;
                  m |= (__ldg(p.key_mask + (long long)t.b * p.mask_bs + tok) != 0 ? 1u : 0u) << (b * KW + c);
                }
          }
          kmask |= (unsigned long long)m << (9 * a);
        }
      }
      // ---- tile start: everyone has left the previous tile (its bos rows / TMEM are free), then stage this one ----
      named_bar_sync(1, 256);
      if (wg == 0) {
        if (X2) {   // learned null key / value (fp32 parameters), the same for every sample
          float* dstk = reinterpret_cast<float*>(sm + OFF_KBOS);
          float* dstv = reinterpret_cast<float*>(sm + OFF_VBOS);
          for (int i = qrow; i < INNER; i += 128) { dstk[i] = __ldg(p.null_k + i); dstv[i] = __ldg(p.null_v + i); }
        } else {                     // bos key / value: sequence row 0 of this sample, 8 heads x 64 channels -> fp32
          const bf16* src = (qrow < 64 ? p.k0 + (long long)t.b * p.k_bs : p.v0 + (long long)t.b * p.v_bs);
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + (qrow & 63));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
          float* dst8 = reinterpret_cast<float*>(sm + (qrow < 64 ? OFF_KBOS : OFF_VBOS)) + (qrow & 63) * 8;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f2 = unpack_bf16x2(w4[e]);
            dst8[2 * e] = f2.x; dst8[2 * e + 1] = f2.y;
          }
        }
      }
      // the bos query (sequence row 0) attends only to itself (nuwa_pytorch.py:608): its output is its value row
      if (!X2 && MODE == 0 && t.f == 0 && tile_row(p, t, 0) == 0 && wg == 1 && qrow < 64) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.v0 + (long long)t.b * p.v_bs) + qrow);
        reinterpret_cast<uint4*>(p.o + (long long)t.b * p.o_bs)[qrow] = v;
      }
      named_bar_sync(1, 256);
      if (lane == 0) mbar_arrive(tready);
      const bool dbg = p.dbg != nullptr && cta == 0 && k == 0 && (warp == 2 || warp == 6) && lane == 0;
      long long* dst_dbg = p.dbg + (warp == 6 ? 32 : 0);
      if (dbg) dst_dbg[0] = clock64();

      if constexpr (PV) {
        // PV mode: this thread's row of dS for its warpgroup's heads -> slot-0 weight (smem, fp32) + bf16 pairs in TMEM at the
        // columns the mix would have left P' in (column 8u + h = slots 1 + 2u, 2 + 2u)
        for (int hp = 0; hp < NH / 2; ++hp) {
          const int h = 2 * hp + wg;
          uint32_t w[NU + 1];
#pragma unroll
          for (int i = 0; i < NU + 1; ++i) w[i] = 0u;
          if (qok) {
            const uint4* src = reinterpret_cast<const uint4*>(p.pv_in + (((long long)t.b * NH + h) * p.nv + vpos) * p.s_jp);
#pragma unroll
            for (int i = 0; i < (NU + 1) / 4; ++i)
              if (8 * i < p.s_jp) {
                const uint4 v4 = __ldg(src + i);
                w[4 * i] = v4.x; w[4 * i + 1] = v4.y; w[4 * i + 2] = v4.z; w[4 * i + 3] = v4.w;
              }
          }
          pbos_s[h * 128 + qrow] = __uint_as_float(w[0] << 16);
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            uint32_t pk[1] = {__funnelshift_r(w[u], w[u + 1], 16)};   // (slot 1 + 2u, slot 2 + 2u)
            tmem_st_x1(tlane + T_P + 8 * u + h, pk);
          }
        }
      }
      // ================= phase 1: band extraction + softmax =================
      for (int hp = 0; hp < (PV ? 0 : NH / 2); ++hp) {
        const int h = 2 * hp + wg;
        float sv[MAXJ + 1];   // scores, then probabilities, of this thread's query for head h: bos, then 9 per frame offset
        {  // bos key (slot 0): 64-term dot product of this thread's Q row (SWIZZLE_128B tile, as the UMMA reads it) with
           // k_bos[h]; runs while the first key tile of the head is still in flight
          const int qb = 2 * (hp & 1) + wg;
          mbar_wait(&qfull[qb], qu & 1);
          const uint8_t* qt = sm + OFF_Q + qb * Q_BYTES + qrow * 128;
          const float4* kb4 = reinterpret_cast<const float4*>(sm + OFF_KBOS) + h * (DH / 4);
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 qv = *reinterpret_cast<const uint4*>(qt + ((c ^ (qrow & 7)) << 4));
            const float4 k0 = kb4[2 * c], k1 = kb4[2 * c + 1];
            const float2 a0 = unpack_bf16x2(qv.x), a1 = unpack_bf16x2(qv.y), a2 = unpack_bf16x2(qv.z), a3 = unpack_bf16x2(qv.w);
            d4[0] = fmaf(a0.x, k0.x, d4[0]); d4[1] = fmaf(a0.y, k0.y, d4[1]);
            d4[2] = fmaf(a1.x, k0.z, d4[2]); d4[3] = fmaf(a1.y, k0.w, d4[3]);
            d4[0] = fmaf(a2.x, k1.x, d4[0]); d4[1] = fmaf(a2.y, k1.y, d4[1]);
            d4[2] = fmaf(a3.x, k1.z, d4[2]); d4[3] = fmaf(a3.y, k1.w, d4[3]);
          }
          sv[0] = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&qempty[qb]);
          if (hp & 1) ++qu;
        }
        // window slots: unit a fills sv[1 + 9a .. 9 + 9a] (kernel row b, column c at 3b + c; rows beyond kh stay masked)
#pragma unroll
        for (int a = 0; a < MAXKT; ++a) {
          if (a < p.kt && ((t.real_mask >> a) & 1)) {
            mbar_wait(&sfull[wg], su & 1);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld_x64(tlane + T_S + wg * 160 + wq * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[wg]);
            ++su;
            // dump the 64-column window to this thread's private row (16 vector stores), pick the 9 in-band entries back
            // with their precomputed offsets: ~40 instructions per unit instead of a test + predicated store per column
#pragma unroll
            for (int i = 0; i < 16; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(drow_u + 16 * i), "r"(v[4 * i]), "r"(v[4 * i + 1]),
                           "r"(v[4 * i + 2]), "r"(v[4 * i + 3]) : "memory");
            const uint32_t km = X2 ? (uint32_t)(kmask >> (9 * a)) & 0x1ffu : 0x1ffu;
#pragma unroll
            for (int e = 0; e < MAXKH * KW; ++e)
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sv[1 + 9 * a + e])
                           : "r"(drow_u + ((!X2 || ((km >> e) & 1u)) ? goff[e] : 256u)) : "memory");
          } else if (a < p.kt && ((t.zero_mask >> a) & 1)) {
            // in-volume key frame beyond the sequence: visible zero keys (score 0), SURVEY D16
#pragma unroll
            for (int e = 0; e < MAXKH * KW; ++e) sv[1 + 9 * a + e] = ((valid9 >> e) & 1u) ? 0.f : -FLT_MAX;
          } else {
#pragma unroll
            for (int e = 0; e < MAXKH * KW; ++e) sv[1 + 9 * a + e] = -FLT_MAX;
          }
        }
        if constexpr (SC) {
          // scores mode: this thread's band of head h -> s_out[b][h][query][:] (fp32, 16-byte stores when the internal
          // slot order is the external one, i.e. kernel height 3)
          if (qok) {
            float* So = p.s_out + (((long long)t.b * NH + h) * p.nv + vpos) * p.s_jp;
            auto cv = [&](float v) { return v == -FLT_MAX ? p.s_masked : v * p.s_scale; };
            if (p.kh == MAXKH && (p.s_jp & 3) == 0 && p.s_jp >= ((1 + 9 * p.kt + 3) & ~3)) {
#pragma unroll
              for (int i4 = 0; i4 < (MAXJ + 3) / 4; ++i4) {
                if (4 * i4 < 1 + 9 * p.kt) {
                  float o4[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int j = 4 * i4 + e;
                    o4[e] = (j < MAXJ && j < 1 + 9 * p.kt) ? cv(sv[j < MAXJ ? j : 0]) : 0.f;
                  }
                  *reinterpret_cast<float4*>(So + 4 * i4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                }
              }
            } else {
              So[0] = cv(sv[0]);
#pragma unroll
              for (int a = 0; a < MAXKT; ++a)
#pragma unroll
                for (int b = 0; b < MAXKH; ++b)
#pragma unroll
                  for (int c = 0; c < KW; ++c)
                    if (a < p.kt && b < p.kh) So[1 + (a * p.kh + b) * KW + c] = cv(sv[1 + 9 * a + 3 * b + c]);
            }
          }
          continue;
        }
        // ---- softmax of this head's row (fp32): bos probability -> smem (fp32), window probabilities -> fp16 pairs in
        //      TMEM, column 8u + h = window entries (2u, 2u + 1) ----
        {
          sv[MAXJ] = -FLT_MAX;
          float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
          for (int i = 0; i < MAXJ; ++i) m4[i & 3] = fmaxf(m4[i & 3], sv[i]);
          const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          const float mneg = -m * p.scale_log2e;
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < MAXJ + 1; ++i) {
            sv[i] = fast_exp2(fmaf(sv[i], p.scale_log2e, mneg));   // masked slots: exp2(-huge) == 0
            l4[i & 3] += sv[i];
          }
          const float inv = 1.0f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
          pbos_s[h * 128 + qrow] = sv[0] * inv;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const __half2 h2 = __floats2half2_rn(sv[1 + 2 * u] * inv, sv[2 + 2 * u] * inv);
            uint32_t pk[1] = {*reinterpret_cast<const uint32_t*>(&h2)};
            tmem_st_x1(tlane + T_P + 8 * u + h, pk);
          }
        }
        if (dbg) dst_dbg[1 + hp] = clock64();
      }
      if constexpr (SC) continue;   // next tile
      // ---- talking heads on the tensor cores: hand P to the MMA thread, convert its fp32 result to bf16 P' in place ----
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (!PV && lane == 0) mbar_arrive(mixgo);
      if (dbg) dst_dbg[5] = clock64();
#pragma unroll 1
      for (int batch = 0; batch < (PV ? 0 : 2); ++batch) {
        mbar_wait(mixfull, mf & 1);
        ++mf;
        tc_fence_after();
        const int u0 = batch ? 20 : 0, u1 = batch ? NU : 20;
#pragma unroll 1
        for (int u = u0 + wg; u < u1; u += 2) {
          uint32_t d[16], o8[8];
          tmem_ld_x16(tlane + T_D + 16 * (u - u0), d);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < NH; ++g) o8[g] = pack_bf16x2(__uint_as_float(d[2 * g]), __uint_as_float(d[2 * g + 1]));
          tmem_st_x8(tlane + T_P + 8 * u, o8);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (batch == 0 && lane == 0) mbar_arrive(mixgo);
      }
      // both warpgroups are done with D: its columns become the A operands / O accumulators.  Zero this warpgroup's A
      // buffer: only the 4-slot-column window of a lane is ever rewritten, the rest of the 160-key row must stay zero
      named_bar_sync(1, 256);
      tc_fence_after();
      {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < 80; c += 16) tmem_st_x16(tlane + T_A + wg * 80 + c, z);
        tmem_st_wait();
      }
      if (dbg) dst_dbg[6] = clock64();

      // ================= phase 2: dense A operand rows + head epilogues =================
      // one unit: the 9 mixed probabilities of frame offset a (halves [9a, 9a + 9) of head g: column 8u + g,
      // u = half / 2) -> this warp's 4-slot-column window of the dense A operand -> publish to the MMA issuer
      auto build_unit = [&](int g, int a) {
        const int h0 = a * 9;
        uint32_t c8[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          uint32_t one[1];
          tmem_ld_x1(tlane + T_P + 8 * ((h0 >> 1) + i) + g, one);
          c8[i] = one[0];
        }
        tmem_ld_wait();
        const uint32_t sh = (uint32_t)(h0 & 1) * 16u;
        uint32_t pr[5];                           // pairs (i0,i1) (i2,i3) (i4,i5) (i6,i7) (i8,.)
#pragma unroll
        for (int i = 0; i < 4; ++i) pr[i] = __funnelshift_r(c8[i], c8[i + 1], sh);
        pr[4] = c8[4] >> sh;
        uint32_t PA[MAXKH], PB[MAXKH];
        PA[0] = pr[0];                       PB[0] = pr[1] & 0xffffu;
        PA[1] = prmt(pr[1], pr[2], 0x5432u); PB[1] = pr[2] >> 16;
        PA[2] = pr[3];                       PB[2] = pr[4] & 0xffffu;
        uint32_t win[32];
#pragma unroll
        for (int scw = 0; scw < 4; ++scw) {
          // kernel row b = scw - r
          uint32_t pa, pb;
          if (scw == 0) { pa = r ? 0u : PA[0]; pb = r ? 0u : PB[0]; }
          else if (scw == 3) { pa = r ? PA[2] : 0u; pb = r ? PB[2] : 0u; }
          else { pa = r ? PA[scw - 1] : PA[scw]; pb = r ? PB[scw - 1] : PB[scw]; }
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) win[scw * 8 + w8] = prmt(pa, pb, sel[w8]);
        }
        mbar_wait(&aempty[wg], (au & 1) ^ 1);
        tc_fence_after();
        tmem_st_x32(tlane + T_A + wg * 80 + wq * 16, win);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[wg]);
        ++au;
      };
      const int a_first = __ffs(t.real_mask) - 1;
      for (int gp = 0; gp < NH / 2; ++gp) {
        const int g = 2 * gp + wg;
#pragma unroll 1
        for (int a = 0; a < p.kt; ++a) {
          if (!((t.real_mask >> a) & 1)) continue;
          if (gp > 0 && a == a_first) continue;      // published before the previous head's epilogue
          build_unit(g, a);
        }
        // the next head's first unit goes out BEFORE this head's epilogue: the issuer starts it as soon as the
        // accumulator has been read (oempty), instead of idling through the epilogue arithmetic and stores
        if (gp + 1 < NH / 2) build_unit(g + 2, a_first);
        // ---- head epilogue: O (fp32, TMEM) + mixed bos probability x bos value -> bf16 -> global ----
        float pbos = 0.f;
        if constexpr (PV) {
          pbos = pbos_s[g * 128 + qrow];
        } else {
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) pbos = fmaf(Wsm[g * NH + hh], pbos_s[hh * 128 + qrow], pbos);
        }
        mbar_wait(&ofull[wg], ou & 1);
        tc_fence_after();
        uint32_t ov[64];
        tmem_ld_x64(tlane + T_O + wg * DH, ov);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&oempty[wg]);
        ++ou;
        if (qok) {
          bf16* dst = p.o + (long long)t.b * p.o_bs + (long long)(p.q_tok0 + vpos) * p.o_rs + g * DH;
#pragma unroll
          for (int c8i = 0; c8i < 8; ++c8i) {
            const float4 va = *reinterpret_cast<const float4*>(vbos + g * DH + c8i * 8);
            const float4 vb4 = *reinterpret_cast<const float4*>(vbos + g * DH + c8i * 8 + 4);
            const float vf[8] = {va.x, va.y, va.z, va.w, vb4.x, vb4.y, vb4.z, vb4.w};
            uint32_t outw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              outw[e] = pack_bf16x2(fmaf(pbos, vf[2 * e], __uint_as_float(ov[c8i * 8 + 2 * e])),
                                    fmaf(pbos, vf[2 * e + 1], __uint_as_float(ov[c8i * 8 + 2 * e + 1])));
            *reinterpret_cast<uint4*>(dst + c8i * 8) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
          }
        }
        if (dbg) dst_dbg[7 + gp] = clock64();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

long long* g_umma_dbg = nullptr;

}  // namespace

// measurement aid (tools/umma_stamps.py; not part of include/nuwa_b200.h): 64 int64 of device memory or NULL
extern "C" void nuwa_debug_umma_stamps(long long* dev_ptr) { g_umma_dbg = dev_ptr; }

namespace {

struct HostMaps {
  CUtensorMap flat, m8, m4, m2;
  int rows5d;
};

// TMA views of one [B][tokens][row_stride] bf16 buffer: the flat token view (zero fill past `tokens`) and, when the row
// dilation divides the grid height, the 5-D view of its complete 16-token grid rows: grid row R = yhi * dh + ylo, so a box
// of n consecutive yhi is n rows spaced by the row dilation.  tok0 = flat token index of grid position 0.
int build_maps(HostMaps& m, const bf16* base, int rs, long long bs, int tokens, int tok0, int dh, int B) {
  const uint64_t dims[3] = {(uint64_t)rs, (uint64_t)tokens, (uint64_t)B};
  const uint64_t strides[3] = {2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  const uint32_t box[3] = {DH, GW, 1};
  int rc = encode_map_bf16_sw128(&m.flat, base, 3, dims, strides, box);
  if (rc != NUWA_OK) return rc;
  m.rows5d = 0;
  if (GW % dh == 0) m.rows5d = ((tokens - tok0) / GW / dh) * dh;
  m.m8 = m.m4 = m.m2 = m.flat;
  if (m.rows5d > 0) {
    const uint64_t d5[5] = {(uint64_t)rs, (uint64_t)GW, (uint64_t)dh, (uint64_t)(m.rows5d / dh), (uint64_t)B};
    const uint64_t s5[5] = {2, (uint64_t)rs * 2, (uint64_t)GW * rs * 2, (uint64_t)dh * GW * rs * 2, (uint64_t)bs * 2};
    const void* g0 = base + (long long)tok0 * rs;
    const uint32_t b8[5] = {DH, GW, 1, 8, 1}, b4[5] = {DH, GW, 1, 4, 1}, b2[5] = {DH, GW, 1, 2, 1};
    if ((rc = encode_map_bf16_sw128(&m.m8, g0, 5, d5, s5, b8)) != NUWA_OK) return rc;
    if ((rc = encode_map_bf16_sw128(&m.m4, g0, 5, d5, s5, b4)) != NUWA_OK) return rc;
    if ((rc = encode_map_bf16_sw128(&m.m2, g0, 5, d5, s5, b2)) != NUWA_OK) return rc;
  }
  return NUWA_OK;
}

int tiles_per_frame(int dh) {
  if (dh >= 4) return (min(dh, GW) + 1) / 2;
  int tiles = 0;
  for (int r = 0; r < dh; ++r) tiles += ((GW - r + dh - 1) / dh + 7) / 8;
  return tiles;
}

int launch_umma(const HostMaps& qm, const HostMaps& km, UmmaArgs& a, cudaStream_t stream) {
  a.q_rows5d = qm.rows5d; a.kv_rows5d = km.rows5d;
  a.tpf = tiles_per_frame(a.dh);
  a.ntiles = a.nf * a.tpf * a.B;
  a.dbg = g_umma_dbg;
  const int grid = min(a.ntiles, device_sm_count());
  auto launch = [&](void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                 const CUtensorMap, const CUtensorMap, const CUtensorMap, const UmmaArgs)) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return NUWA_ERR_CUDA;
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(qm.flat, qm.m8, qm.m4, qm.m2, km.flat, km.m8, km.m4, km.m2, a);
    return NUWA_OK;
  };
  int lrc;
  const int mode = a.s_out != nullptr ? 1 : (a.pv_in != nullptr ? 2 : 0);
#define NUWA_UMMA_DW(X2V, MODEV)                                                         \
  (a.dw == 1 ? launch(attn_3dna_umma_kernel<1, X2V, MODEV>)                             \
             : (a.dw == 2 ? launch(attn_3dna_umma_kernel<2, X2V, MODEV>) : launch(attn_3dna_umma_kernel<4, X2V, MODEV>)))
  if (a.abs_frames) lrc = mode == 1 ? NUWA_UMMA_DW(true, 1) : (mode == 2 ? NUWA_UMMA_DW(true, 2) : NUWA_UMMA_DW(true, 0));
  else lrc = mode == 1 ? NUWA_UMMA_DW(false, 1) : (mode == 2 ? NUWA_UMMA_DW(false, 2) : NUWA_UMMA_DW(false, 0));
#undef NUWA_UMMA_DW
  if (lrc != NUWA_OK) return lrc;
  NUWA_CHECK_LAUNCH();
  return NUWA_OK;
}

}  // namespace

// Envelope: full pass (t0 == 0, nq == nv + 1) over a 16-wide token grid, H == 8, dh == 64, kw == 3, kh <= 3, kt <= 5,
// column dilation 1 / 2 / 4, q|k|v rows sharing one token stride; causal or centred windows.  NUWA_ERR_INVALID outside
// it (nothing launched; the caller falls back to the halo / gather kernels).
int attn_3dna_umma(const AttnParams& p, cudaStream_t stream) {
  if (p.fmap != GW || p.t0 != 0 || p.t0_ptr != nullptr) return NUWA_ERR_INVALID;
  if (p.H != NH || p.dh != DH || p.nq != p.nv + 1 || p.nv <= 0 || p.B <= 0) return NUWA_ERR_INVALID;
  if (p.kw != KW || p.kh < 1 || p.kh > MAXKH || p.kt < 1 || p.kt > MAXKT) return NUWA_ERR_INVALID;
  if (p.dt <= 0 || p.dh_ <= 0 || !(p.dw == 1 || p.dw == 2 || p.dw == 4)) return NUWA_ERR_INVALID;
  if (!(p.kh & 1) || !(p.kt & 1)) return NUWA_ERR_INVALID;  // odd kernels (nuwa_pytorch.py:410)
  if (p.max_frames <= 0 || p.nv > p.max_frames * GW * GW) return NUWA_ERR_INVALID;
  if (p.head_scale != nullptr || p.bias != nullptr || p.key_mask != nullptr || p.null_k != nullptr) return NUWA_ERR_INVALID;
  const bf16* q = reinterpret_cast<const bf16*>(p.q);
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* v = reinterpret_cast<const bf16*>(p.v);
  const long long koff = k - q, voff = v - q;
  if (p.k_rs != p.q_rs || p.v_rs != p.q_rs || p.k_bs != p.q_bs || p.v_bs != p.q_bs) return NUWA_ERR_INVALID;
  if (koff < 0 || voff < 0 || koff + INNER > p.q_rs || voff + INNER > p.q_rs) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.q_bs % 8) || (koff % 8) || (voff % 8) || (p.o_rs % 8) || (p.o_bs % 8)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.o) & 15)) return NUWA_ERR_INVALID;

  HostMaps m;   // q, k and v live in one buffer: one map set serves both sides (sequence row 0 = bos, grid from row 1)
  int rc = build_maps(m, q, p.q_rs, p.q_bs, p.nv + 1, 1, p.dh_, p.B);
  if (rc != NUWA_OK) return rc;

  UmmaArgs a = {};
  a.B = p.B; a.nv = p.nv;
  a.nf = (p.nv + GW * GW - 1) / (GW * GW);
  a.maxf = p.max_frames;
  a.kt = p.kt; a.kh = p.kh; a.dt = p.dt; a.dh = p.dh_; a.dw = p.dw; a.causal = p.causal;
  a.koff = (int)koff; a.voff = (int)voff;
  a.q_tok0 = a.kv_tok0 = 1;
  a.scale_log2e = p.qscale * 1.4426950408889634f;
  a.talk = p.talk;
  a.o = reinterpret_cast<bf16*>(p.o); a.o_bs = p.o_bs; a.o_rs = p.o_rs;
  a.k0 = k; a.v0 = v; a.k_bs = p.k_bs; a.v_bs = p.v_bs;
  return launch_umma(m, m, a, stream);
}

// Backward scores of Sparse3DNA on the same kernel (phase 1 only, twice): S[b][h][q][j] = qscale q . k_j (masked slots
// -FLT_MAX, visible zero keys 0) and dP'[b][h][q][j] = dO . v_j (0 in both cases), fp32, row pitch jp -- what
// attn3dna_bwd_scores (gather kernel) writes.  `p` follows the backward convention: p.q = first non-bos query row, p.k / p.v =
// row 0 (bos) of the same q|k|v buffer, p.nq = p.nv = number of video tokens, p.t0 == 1.  Same envelope as attn_3dna_umma.
int attn_3dna_umma_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                          cudaStream_t stream) {
  if (p.fmap != GW || p.t0 != 1 || p.t0_ptr != nullptr || S == nullptr || dPp == nullptr || dO == nullptr) return NUWA_ERR_INVALID;
  if (p.H != NH || p.dh != DH || p.nq != p.nv || p.nv <= 0 || p.B <= 0) return NUWA_ERR_INVALID;
  if (p.kw != KW || p.kh < 1 || p.kh > MAXKH || p.kt < 1 || p.kt > MAXKT) return NUWA_ERR_INVALID;
  if (p.dt <= 0 || p.dh_ <= 0 || !(p.dw == 1 || p.dw == 2 || p.dw == 4)) return NUWA_ERR_INVALID;
  if (!(p.kh & 1) || !(p.kt & 1) || jp < 1 + p.kt * p.kh * KW) return NUWA_ERR_INVALID;
  if (p.max_frames <= 0 || p.nv > p.max_frames * GW * GW) return NUWA_ERR_INVALID;
  if (p.head_scale != nullptr || p.bias != nullptr || p.key_mask != nullptr || p.null_k != nullptr) return NUWA_ERR_INVALID;
  const bf16* q1 = reinterpret_cast<const bf16*>(p.q);       // sequence row 1
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* v = reinterpret_cast<const bf16*>(p.v);
  const bf16* row0 = q1 - p.q_rs;                              // q of the bos row: start of the token rows
  const long long koff = k - row0, voff = v - row0;
  if (p.k_rs != p.q_rs || p.v_rs != p.q_rs || p.k_bs != p.q_bs || p.v_bs != p.q_bs) return NUWA_ERR_INVALID;
  if (koff < 0 || voff < 0 || koff + INNER > p.q_rs || voff + INNER > p.q_rs) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.q_bs % 8) || (koff % 8) || (voff % 8) || (do_rs % 8) || (do_bs % 8) || do_rs < INNER) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(row0) & 15) || (reinterpret_cast<uintptr_t>(dO) & 15) || (reinterpret_cast<uintptr_t>(S) & 15) ||
      (reinterpret_cast<uintptr_t>(dPp) & 15))
    return NUWA_ERR_INVALID;

  HostMaps qm, dm, km;
  int rc = build_maps(km, row0, p.q_rs, p.q_bs, p.nv + 1, 1, p.dh_, p.B);               // keys / values: bos row first
  if (rc != NUWA_OK) return rc;
  if ((rc = build_maps(qm, q1, p.q_rs, p.q_bs, p.nv, 0, p.dh_, p.B)) != NUWA_OK) return rc;   // queries from row 1
  if ((rc = build_maps(dm, reinterpret_cast<const bf16*>(dO), do_rs, do_bs, p.nv, 0, p.dh_, p.B)) != NUWA_OK) return rc;

  UmmaArgs a = {};
  a.B = p.B; a.nv = p.nv;
  a.nf = (p.nv + GW * GW - 1) / (GW * GW);
  a.maxf = p.max_frames;
  a.kt = p.kt; a.kh = p.kh; a.dt = p.dt; a.dh = p.dh_; a.dw = p.dw; a.causal = p.causal;
  a.q_tok0 = 0; a.kv_tok0 = 1;
  a.scale_log2e = 0.f;
  a.k_bs = p.k_bs; a.v_bs = p.v_bs;
  a.s_jp = jp;
  // logits: Q = q, K = k
  a.koff = (int)koff; a.voff = (int)koff;
  a.k0 = k; a.v0 = k;
  a.s_out = S; a.s_scale = p.qscale; a.s_masked = -FLT_MAX;
  if ((rc = launch_umma(qm, km, a, stream)) != NUWA_OK) return rc;
  // dP' = dO V^T: Q = dO, K = v
  a.koff = (int)voff; a.voff = (int)voff;
  a.k0 = v; a.v0 = v;
  a.s_out = dPp; a.s_scale = 1.0f; a.s_masked = 0.0f;
  return launch_umma(dm, km, a, stream);
}

// dq of the Sparse3DNA backward on the same kernel in PV mode: dq[q] = sum_j dS[q][j] k_j over [bos | window] with dS (bf16,
// [B][H][nq][jp], already multiplied by dh^-0.5) in the place of the probabilities and V := K.  `p` in the backward convention
// (see attn_3dna_umma_scores); dq rows of the non-bos queries (dq_bs / dq_rs).  Kernel height 3 only (internal slot order ==
// external), jp % 8 == 0; NUWA_ERR_INVALID otherwise (the gather kernel nuwa_attn3dna_bwd_dq takes those).
int attn_3dna_umma_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t stream) {
  if (p.fmap != GW || p.t0 != 1 || p.t0_ptr != nullptr || dS == nullptr || dq == nullptr) return NUWA_ERR_INVALID;
  if (p.H != NH || p.dh != DH || p.nq != p.nv || p.nv <= 0 || p.B <= 0) return NUWA_ERR_INVALID;
  if (p.kw != KW || p.kh != MAXKH || p.kt < 1 || p.kt > MAXKT || !(p.kt & 1)) return NUWA_ERR_INVALID;
  if (p.dt <= 0 || p.dh_ <= 0 || !(p.dw == 1 || p.dw == 2 || p.dw == 4)) return NUWA_ERR_INVALID;
  if (jp < 1 + p.kt * MAXKH * KW || (jp % 8) || p.max_frames <= 0 || p.nv > p.max_frames * GW * GW) return NUWA_ERR_INVALID;
  if (p.head_scale != nullptr || p.bias != nullptr || p.key_mask != nullptr || p.null_k != nullptr) return NUWA_ERR_INVALID;
  const bf16* q1 = reinterpret_cast<const bf16*>(p.q);
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* row0 = q1 - p.q_rs;
  const long long koff = k - row0;
  if (p.k_rs != p.q_rs || p.k_bs != p.q_bs || koff < 0 || koff + INNER > p.q_rs) return NUWA_ERR_INVALID;
  if ((p.q_rs % 8) || (p.q_bs % 8) || (koff % 8) || (dq_rs % 8) || (dq_bs % 8)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(row0) & 15) || (reinterpret_cast<uintptr_t>(dS) & 15) || (reinterpret_cast<uintptr_t>(dq) & 15))
    return NUWA_ERR_INVALID;
  HostMaps km;
  const int rc = build_maps(km, row0, p.q_rs, p.q_bs, p.nv + 1, 1, p.dh_, p.B);
  if (rc != NUWA_OK) return rc;
  UmmaArgs a = {};
  a.B = p.B; a.nv = p.nv;
  a.nf = (p.nv + GW * GW - 1) / (GW * GW);
  a.maxf = p.max_frames;
  a.kt = p.kt; a.kh = p.kh; a.dt = p.dt; a.dh = p.dh_; a.dw = p.dw; a.causal = p.causal;
  a.koff = a.voff = (int)koff;          // V := K
  a.q_tok0 = 0; a.kv_tok0 = 1;
  a.k0 = k; a.v0 = k; a.k_bs = p.k_bs; a.v_bs = p.k_bs;
  a.o = reinterpret_cast<bf16*>(dq); a.o_bs = dq_bs; a.o_rs = dq_rs;
  a.pv_in = reinterpret_cast<const bf16*>(dS); a.s_jp = jp;
  return launch_umma(km, km, a, stream);
}

// SparseCross2DNA (nuwa_pytorch.py:851-895) on the same kernel: the nq queries at video positions 0 .. nq-1 (p.q / p.o
// point at the first of them, i.e. past the bos row; p.t0 == 1) each see slot 0 = the learned null key / value and the
// centred ck x ck window, dilation cdil, at their own grid position in every context frame, under the context mask.
// Envelope: 16-wide grid, H == 8, dh == 64, ck == 3, cdil 1 / 2 / 4, <= 5 context frames, k|v rows sharing one stride.
static int cross2dna_setup(const AttnParams& p, HostMaps& qm, HostMaps& km, UmmaArgs& a) {
  if (p.fmap != GW || p.t0 != 1 || p.t0_ptr != nullptr) return NUWA_ERR_INVALID;
  if (p.H != NH || p.dh != DH || p.nq <= 0 || p.B <= 0) return NUWA_ERR_INVALID;
  if (p.ck != KW || !(p.cdil == 1 || p.cdil == 2 || p.cdil == 4)) return NUWA_ERR_INVALID;
  const int frames = (p.jmax - 1) / (KW * KW);
  if (frames < 1 || frames > MAXKT || p.jmax != 1 + frames * KW * KW) return NUWA_ERR_INVALID;
  if (p.head_scale != nullptr || p.bias != nullptr || p.null_k == nullptr || p.null_v == nullptr) return NUWA_ERR_INVALID;
  if (p.key_mask != nullptr && p.mask_bs < frames * GW * GW) return NUWA_ERR_INVALID;
  const bf16* q = reinterpret_cast<const bf16*>(p.q);
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* v = reinterpret_cast<const bf16*>(p.v);
  const bf16* kvbase = k < v ? k : v;
  const long long koff = k - kvbase, voff = v - kvbase;
  if (p.k_rs != p.v_rs || p.k_bs != p.v_bs || koff + INNER > p.k_rs || voff + INNER > p.k_rs) return NUWA_ERR_INVALID;
  if (INNER > p.q_rs || (p.q_rs % 8) || (p.q_bs % 8) || (p.k_rs % 8) || (p.k_bs % 8) || (koff % 8) || (voff % 8))
    return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(kvbase) & 15)) return NUWA_ERR_INVALID;

  int rc = build_maps(qm, q, p.q_rs, p.q_bs, p.nq, 0, p.cdil, p.B);
  if (rc != NUWA_OK) return rc;
  if ((rc = build_maps(km, kvbase, p.k_rs, p.k_bs, frames * GW * GW, 0, p.cdil, p.B)) != NUWA_OK) return rc;

  a = UmmaArgs{};
  a.B = p.B; a.nv = p.nq;
  a.nf = (p.nq + GW * GW - 1) / (GW * GW);
  a.maxf = a.nf;
  a.kt = frames; a.kh = KW; a.dt = 1; a.dh = p.cdil; a.dw = p.cdil; a.causal = 0;
  a.koff = (int)koff; a.voff = (int)voff;
  a.abs_frames = 1;
  a.null_k = p.null_k; a.null_v = p.null_v;
  a.key_mask = p.key_mask; a.mask_bs = p.mask_bs;
  a.scale_log2e = p.qscale * 1.4426950408889634f;
  a.talk = p.talk;
  a.k0 = k; a.v0 = v; a.k_bs = p.k_bs; a.v_bs = p.v_bs;
  return NUWA_OK;
}

int attn_cross2dna_umma(const AttnParams& p, cudaStream_t stream) {
  if ((p.o_rs % 8) || (p.o_bs % 8) || (reinterpret_cast<uintptr_t>(p.o) & 15)) return NUWA_ERR_INVALID;
  HostMaps qm, km;
  UmmaArgs a;
  const int rc = cross2dna_setup(p, qm, km, a);
  if (rc != NUWA_OK) return rc;
  a.o = reinterpret_cast<bf16*>(p.o); a.o_bs = p.o_bs; a.o_rs = p.o_rs;
  return launch_umma(qm, km, a, stream);
}

// dq of the SparseCross2DNA backward (non-bos queries) in PV mode, as attn_3dna_umma_dq: slot 0 multiplies the learned null key
int attn_cross2dna_umma_dq(const AttnParams& p, const void* dS, int jp, void* dq, long long dq_bs, int dq_rs, cudaStream_t stream) {
  if (dS == nullptr || dq == nullptr || jp < p.jmax || (jp % 8) || (dq_rs % 8) || (dq_bs % 8)) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(dS) & 15) || (reinterpret_cast<uintptr_t>(dq) & 15)) return NUWA_ERR_INVALID;
  HostMaps qm, km;
  UmmaArgs a;
  const int rc = cross2dna_setup(p, qm, km, a);
  if (rc != NUWA_OK) return rc;
  a.voff = a.koff;                       // V := K
  a.v0 = a.k0;
  a.null_v = p.null_k;                   // slot 0: dS[0] * null key
  a.key_mask = nullptr;                  // masked slots carry dS == 0
  a.o = reinterpret_cast<bf16*>(dq); a.o_bs = dq_bs; a.o_rs = dq_rs;
  a.pv_in = reinterpret_cast<const bf16*>(dS); a.s_jp = jp;
  return launch_umma(km, km, a, stream);
}

// Backward scores of SparseCross2DNA (non-bos queries) in scores mode, as attn_3dna_umma_scores: S = qscale q . k_j over
// [null key | window] with masked context tokens at -FLT_MAX, dP' = dO . v_j (null value in slot 0, 0 where masked).
int attn_cross2dna_umma_scores(const AttnParams& p, const void* dO, long long do_bs, int do_rs, float* S, float* dPp, int jp,
                               cudaStream_t stream) {
  if (S == nullptr || dPp == nullptr || dO == nullptr || jp < p.jmax) return NUWA_ERR_INVALID;
  if ((do_rs % 8) || (do_bs % 8) || do_rs < INNER) return NUWA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(dO) & 15) || (reinterpret_cast<uintptr_t>(S) & 15) || (reinterpret_cast<uintptr_t>(dPp) & 15))
    return NUWA_ERR_INVALID;
  HostMaps qm, km, dm;
  UmmaArgs a;
  int rc = cross2dna_setup(p, qm, km, a);
  if (rc != NUWA_OK) return rc;
  if ((rc = build_maps(dm, reinterpret_cast<const bf16*>(dO), do_rs, do_bs, p.nq, 0, p.cdil, p.B)) != NUWA_OK) return rc;
  a.s_jp = jp;
  a.voff = a.koff;
  a.s_out = S; a.s_scale = p.qscale; a.s_masked = -FLT_MAX;
  if ((rc = launch_umma(qm, km, a, stream)) != NUWA_OK) return rc;
  const bf16* k = reinterpret_cast<const bf16*>(p.k);
  const bf16* v = reinterpret_cast<const bf16*>(p.v);
  a.koff = a.voff = (int)(v - (k < v ? k : v));
  a.null_k = p.null_v;                 // slot 0 of dP' = dO . null_v
  a.s_out = dPp; a.s_scale = 1.0f; a.s_masked = 0.0f;
  return launch_umma(dm, km, a, stream);
}

}  // namespace nuwa
