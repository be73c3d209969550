// Dense scaled-dot-product attention for the instance-embedding block: a fused
// QK^T / softmax / PV kernel on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged in shared memory by the TMA bulk-copy engine).  Replaces
//   /root/reference/code/lib/archs/modules/utils.py:305-329  ScaledDotProductAttention.forward
//     attn = softmax(masked_fill(q k^T / T, -inf), dim=2);  out = attn v
// (the L x L score matrix is never materialised; `attn` is produced by a separate kernel only
// when the caller asks for it).
//
// Head width in this model is d_k = d_v = 12 (modules/config.py:22-25): far too thin to make the
// tensor pipe the limiter -- per 128x128 tile the MMAs need ~0.4k cycles, the 16384 exponentials
// ~1k cycles of MUFU.  So the spare tensor throughput is spent on accuracy: every operand is split
// into bf16 hi + lo parts and each product is computed as hi*hi + hi*lo + lo*hi (fp32 accumulate
// in TMEM), which keeps the result within ~1e-5 of the fp32 reference instead of bf16's ~1e-2.
//
// Pipeline (one CTA = 128 queries of one head, 2 CTAs per SM so that one CTA's softmax overlaps
// the other's MMAs):
//   prep kernel   q,k,v fp32 -> bf16 hi/lo tiles already in the canonical UMMA K-major
//                 (no-swizzle, 8x16B core matrix) byte order, q pre-scaled by log2(e)/T
//   warp 0        producer: one cp.async.bulk (TMA) of 16 KB per key tile [Kh|Kl|Vh^T|Vl^T]
//   warp 1        tcgen05.alloc; one elected lane issues S = Q K^T (3 MMAs, N=128) and
//                 O += P V (24 MMAs, N=16); completion via tcgen05.commit -> mbarrier
//   warps 2..5    one query row per thread: tcgen05.ld of S, online softmax in the exp2 domain,
//                 P (bf16 hi/lo) -> shared memory as the next A operand, O rescale in TMEM,
//                 epilogue O / l and the log-sum-exp for backward
// Backward: two more tcgen05 kernels with the same structure, recomputing P from the saved
// log-sum-exp:  attn_bwd_dq (rows = queries:  S, dP -> dA = P (dP - delta) -> dQ += dA K) and
// attn_bwd_dkdv (rows = keys: S^T, dP^T -> P^T, dA^T -> dV += P^T dO, dK += dA^T Q); no atomics.
#include "isa_common.cuh"
#include "isa_ptx.cuh"
#include "isa_tcgen05.cuh"
#include <cuda_bf16.h>
#include <math.h>

namespace {

constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr int kDP = 16;            // padded head width
constexpr int kOperandBytes = kTileQ * kDP * 2;     // 4096: one bf16 [128][16] operand tile
constexpr int kKvTileBytes = 4 * kOperandBytes;     // Kh | Kl | Vh^T | Vl^T
constexpr int kQTileBytes = 2 * kOperandBytes;      // Qh | Ql
constexpr int kFwdThreads = 192;
constexpr int kTmemCols = 256;                      // S / P: [0,128), O: [128,160) = [P V_hi | P V_lo]
constexpr int kFwdStages = 4;
// The tensor core adds into its fp32 accumulator with truncation, so thousands of same-sign accumulations drift
// (measured -9e-4 relative at L = 131 072 keys).  Every kFlushTiles key tiles the PV accumulator is therefore folded
// into fp32 registers (round-to-nearest adds) and restarted.
constexpr int kFlushTiles = 16;
constexpr int kBwdThreads = 384;                    // warp 0 producer, warps 1-2 MMA issuers (even / odd sub-tiles), warp 3 idle, warps 4..11 math
constexpr int kBwdTmemCols = 512;
constexpr int kTmemO = 128;

// canonical K-major no-swizzle layouts (units: bytes).  Element (row r, k):
//   (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2     (core matrix = 8 rows x 16 B, contiguous)
__host__ __device__ inline int kmajor_off(int r, int k, int sbo) { return (r >> 3) * sbo + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }
constexpr int kSboQK = 256;    // [128 rows][16 k]: 2 core matrices along K
constexpr int kSboP = 2048;    // [128 rows][128 k]: 16 core matrices along K (also V^T: [16 rows][128 k])

// 8 consecutive fp32 values -> one 16-byte group of the hi operand and one of the lo operand
__device__ __forceinline__ void store_group(unsigned char* hi_base, unsigned char* lo_base, int off, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(x[2 * e], x[2 * e + 1], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------- operand tile writers (prep)
// [128 rows][16 k] K-major operand from src[row][c] (row stride ld, c < width), rows >= n_rows and c >= width zero.
// 256 threads: thread -> (row r = tid/2, k group kg = tid%2)
__device__ __forceinline__ void write_row_tile(unsigned char* dst_hi, unsigned char* dst_lo, const float* __restrict__ src, int ld, int width,
                                               int row0, int n_rows, float scale) {
  const int r = threadIdx.x >> 1, kg = threadIdx.x & 1;
  const int row = row0 + r;
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = kg * 8 + e;
    x[e] = (row < n_rows && c < width) ? __ldg(src + (size_t)row * ld + c) * scale : 0.f;
  }
  store_group(dst_hi, dst_lo, kmajor_off(r, kg * 8, kSboQK), x);
}
// [16 rows = channel][128 k = position] K-major operand (the transpose), 256 threads: thread -> (n = tid/16, k group = tid%16)
__device__ __forceinline__ void write_col_tile(unsigned char* dst_hi, unsigned char* dst_lo, const float* __restrict__ src, int ld, int width,
                                               int row0, int n_rows, float scale) {
  const int n = threadIdx.x >> 4, kgv = threadIdx.x & 15;
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int row = row0 + kgv * 8 + e;
    x[e] = (row < n_rows && n < width) ? __ldg(src + (size_t)row * ld + n) * scale : 0.f;
  }
  store_group(dst_hi, dst_lo, kmajor_off(n, kgv * 8, kSboP), x);
}

struct PrepParams {
  const float* q; const float* k; const float* v;
  unsigned char* qblk;   // [BH][nQt][Qh 4K | Ql 4K]
  unsigned char* kvblk;  // [BH][nKt][Kh | Kl | Vh^T | Vl^T]
  int BH, Lq, Lk, d, dv, nQt, nKt;
  float qscale;          // log2(e) / temperature
};

__global__ void __launch_bounds__(256) attn_prep_kernel(const PrepParams p) {
  const int bh = blockIdx.y, tile = blockIdx.x;
  if (tile < p.nQt) {
    unsigned char* dst = p.qblk + ((size_t)bh * p.nQt + tile) * kQTileBytes;
    write_row_tile(dst, dst + kOperandBytes, p.q + (size_t)bh * p.Lq * p.d, p.d, p.d, tile * kTileQ, p.Lq, p.qscale);
  }
  if (tile < p.nKt) {
    unsigned char* dst = p.kvblk + ((size_t)bh * p.nKt + tile) * kKvTileBytes;
    write_row_tile(dst, dst + kOperandBytes, p.k + (size_t)bh * p.Lk * p.d, p.d, p.d, tile * kTileK, p.Lk, 1.f);
    write_col_tile(dst + 2 * kOperandBytes, dst + 3 * kOperandBytes, p.v + (size_t)bh * p.Lk * p.dv, p.dv, p.dv, tile * kTileK, p.Lk, 1.f);
  }
}

// ---------------------------------------------------------------------------- masks
struct MaskArgs {
  const unsigned char* key_mask;   // [n_mask_rows][Lk] 1 = masked, or null
  const unsigned char* full_mask;  // [n_mask_rows][Lq][Lk] or null
  int n_mask_rows;
};
// ---------------------------------------------------------------------------- attention-probability dropout
// nn.Dropout(attn_dropout) on the softmax probabilities (/root/reference/code/lib/archs/modules/utils.py:311,326), fused:
// the keep / drop decision of element (head-batch bh, query q, key k) is a pure function of (seed, bh, q, k), so the
// forward never stores a mask and the two backward kernels regenerate it.  One 32-bit hash serves the key pair
// (k & ~1, k | 1): 16 bits each against thr = round(p * 65536); kept probabilities are scaled by 1 / (1 - p).
struct DropArgs {
  uint32_t seed;      // already mixed with the call's seed / offset
  uint32_t thr;       // 0 = no dropout
  float scale;        // 1 / (1 - p)
};
__device__ __forceinline__ uint32_t hash32(uint32_t x) {          // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_row_key(const DropArgs& d, int bh, int q) {
  return hash32(d.seed + (uint32_t)bh * 0x9E3779B1u + (uint32_t)q * 0x85EBCA77u);
}
// 16-bit lottery numbers of keys 2*kpair (low half) and 2*kpair + 1 (high half) for the query behind `rowkey`
__device__ __forceinline__ uint32_t drop_pair(uint32_t rowkey, int kpair) { return hash32(rowkey + (uint32_t)kpair * 0xC2B2AE3Du); }
__device__ __forceinline__ bool drop_one(const DropArgs& d, uint32_t rowkey, int k) {
  const uint32_t h = drop_pair(rowkey, k >> 1);
  return ((k & 1) ? (h >> 16) : (h & 0xffffu)) < d.thr;
}

// 128-bit mask of the keys [key0, key0+128) that are masked for query `row` (bit set = masked), incl. keys >= Lk
__device__ __forceinline__ void key_bits(const MaskArgs& m, int bh, int row, int Lq, int Lk, int key0, uint32_t (&bits)[4]) {
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint32_t b = 0;
    const int kb = key0 + w * 32;
    if (kb + 32 > Lk) b = (kb >= Lk) ? 0xffffffffu : (0xffffffffu << (Lk - kb));
    bits[w] = b;
  }
  if (m.key_mask) {
    const unsigned char* km = m.key_mask + (size_t)(bh % m.n_mask_rows) * Lk;
    for (int i = 0; i < 128 && key0 + i < Lk; ++i)
      if (km[key0 + i]) bits[i >> 5] |= 1u << (i & 31);
  }
  if (m.full_mask && row < Lq) {
    const unsigned char* fm = m.full_mask + ((size_t)(bh % m.n_mask_rows) * Lq + row) * Lk;
    for (int i = 0; i < 128 && key0 + i < Lk; ++i)
      if (fm[key0 + i]) bits[i >> 5] |= 1u << (i & 31);
  }
}
// does this key tile need per-element masking at all?  (warp-uniform answer)
__device__ __forceinline__ bool tile_needs_mask(const MaskArgs& m, int bh, int Lk, int key0, int lane) {
  if (m.full_mask) return true;
  bool any = key0 + kTileK > Lk;
  if (m.key_mask) {
    const unsigned char* km = m.key_mask + (size_t)(bh % m.n_mask_rows) * Lk;
    for (int i = lane; i < 128; i += 32)
      if (key0 + i < Lk && km[key0 + i]) any = true;
  }
  return __any_sync(0xffffffffu, any);
}

// ---------------------------------------------------------------------------- forward
struct FwdParams {
  const unsigned char* qblk;
  const unsigned char* kvblk;
  MaskArgs mask;
  DropArgs drop;
  float* out;    // [BH][Lq][dv]
  float* lse2;   // [BH][Lq]  log2-domain log-sum-exp of the scaled scores (for backward)
  int BH, Lq, Lk, dv, nQt, nKt;
};

struct FwdSmem {
  unsigned char q[kQTileBytes];
  unsigned char kv[kFwdStages][kKvTileBytes];
  uint64_t q_full, s_full, p_full, o_full;
  uint64_t kv_full[kFwdStages], kv_empty[kFwdStages];
  uint32_t tmem_base;
};

template <bool MASKED>
__device__ __forceinline__ float fwd_tile_max(uint32_t t_row, const uint32_t (&bits)[4]) {
  float m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float s[32];
    tmem_ld32(t_row + c * 32, s);
    if (MASKED) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (!((bits[c] >> i) & 1u)) m = fmaxf(m, s[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) m = fmaxf(m, fmaxf(s[i], s[i + 1]));
    }
  }
  return m;
}
// exp pass: P = exp2(S - m) goes back into TENSOR MEMORY as the A operand of the PV MMAs, written over the 32 S
// columns this thread has just read: [P hi: 16 packed columns | P lo: 16 packed columns] per 32-key chunk
template <bool MASKED, bool DROP>
__device__ __forceinline__ float fwd_tile_exp(uint32_t t_row, const uint32_t (&bits)[4], float m_safe, const DropArgs& drop, uint32_t rowkey, int key0) {
  float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t sr[32];
    tmem_ld32_issue(t_row + c * 32, sr);
    tmem_ld32_wait(sr);
    uint32_t ph[16], pl[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float p0 = ex2_approx(__uint_as_float(sr[i]) - m_safe);
      float p1 = ex2_approx(__uint_as_float(sr[i + 1]) - m_safe);
      if (MASKED) {
        if ((bits[c] >> i) & 1u) p0 = 0.f;
        if ((bits[c] >> (i + 1)) & 1u) p1 = 0.f;
      }
      l0 += p0;      // the softmax normaliser sums the probabilities BEFORE dropout
      l1 += p1;
      if (DROP) {
        const uint32_t h = drop_pair(rowkey, (key0 + c * 32 + i) >> 1);
        p0 = ((h & 0xffffu) < drop.thr) ? 0.f : p0 * drop.scale;
        p1 = ((h >> 16) < drop.thr) ? 0.f : p1 * drop.scale;
      }
      split2(p0, p1, ph[i >> 1], pl[i >> 1]);
    }
    tmem_st16_issue(t_row + c * 32, ph);
    tmem_st16_issue(t_row + c * 32 + 16, pl);
  }
  tmem_st_wait();
  return l0 + l1;
}

__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_kernel(const FwdParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int nKt = prm.nKt;

  if (threadIdx.x == 0) {
    mbar_init(&sm.q_full, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, 128);
    mbar_init(&sm.o_full, 1);
    for (int s = 0; s < kFwdStages; ++s) { mbar_init(&sm.kv_full[s], 1); mbar_init(&sm.kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ===================== producer: TMA bulk copies =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&sm.q_full, kQTileBytes);
      tma_bulk_g2s(sm.q, prm.qblk + ((size_t)bh * prm.nQt + qt) * kQTileBytes, kQTileBytes, &sm.q_full);
      for (int j = 0; j < nKt; ++j) {
        const int st = j % kFwdStages;
        mbar_wait(&sm.kv_empty[st], ((j / kFwdStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&sm.kv_full[st], kKvTileBytes);
        tma_bulk_g2s(sm.kv[st], prm.kvblk + ((size_t)bh * nKt + j) * kKvTileBytes, kKvTileBytes, &sm.kv_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform code, one elected lane executes the instructions =====================
    {
      const uint32_t el = elect_one();
      constexpr uint32_t idesc_s = make_idesc(kTileQ, kTileK);
      constexpr uint32_t idesc_o2 = make_idesc(kTileQ, 2 * kDP);   // P_hi x [V_hi ; V_lo] -> 32 columns
      constexpr uint32_t idesc_o1 = make_idesc(kTileQ, kDP);
      const uint64_t q_hi = make_desc(smem_u32(sm.q), 128, kSboQK), q_lo = q_hi + (kOperandBytes >> 4);
      const uint64_t k0 = make_desc(smem_u32(sm.kv[0]), 128, kSboQK);      // K operands of stage 0
      const uint64_t v0 = make_desc(smem_u32(sm.kv[0]) + 2 * kOperandBytes, 128, kSboP);   // V^T operands of stage 0
      mbar_wait(&sm.q_full, 0);
      for (int j = 0; j < nKt; ++j) {
        const int st = j % kFwdStages;
        const uint64_t kd = k0 + (uint32_t)st * (kKvTileBytes >> 4), vd = v0 + (uint32_t)st * (kKvTileBytes >> 4);
        mbar_wait(&sm.kv_full[st], (j / kFwdStages) & 1);
        tc_fence_after();
        // S = Qh Kh^T + Qh Kl^T + Ql Kh^T
        umma_bf16_e(el, tmem, q_hi, kd, idesc_s, 0);
        umma_bf16_e(el, tmem, q_hi, kd + (kOperandBytes >> 4), idesc_s, 1);
        umma_bf16_e(el, tmem, q_lo, kd, idesc_s, 1);
        umma_commit_e(el, &sm.s_full);
        // O += Ph [Vh ; Vl] + Pl Vh, P read from tensor memory (written over S by the softmax warps)
        mbar_wait(&sm.p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kTileK / 16; ++kk) {
          const uint32_t ca = (kk >> 1) * 32 + (kk & 1) * 8;
          const uint32_t ko = (kk * 256) >> 4;
          umma_bf16_ts_e(el, tmem + kTmemO, tmem + ca, vd + ko, idesc_o2, (j % kFlushTiles != 0 || kk > 0) ? 1u : 0u);
          umma_bf16_ts_e(el, tmem + kTmemO, tmem + ca + 16, vd + ko, idesc_o1, 1);
        }
        umma_commit_e(el, &sm.kv_empty[st]);
      }
      umma_commit_e(el, &sm.o_full);
    }
  } else {
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch
    const int r = quarter * 32 + lane;            // row inside the tile
    const int row = qt * kTileQ + r;
    const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    float o_reg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o_reg[i] = 0.f;
    for (int j = 0; j < nKt; ++j) {
      const int key0 = j * kTileK;
      const bool masked = tile_needs_mask(prm.mask, bh, prm.Lk, key0, lane);   // warp-uniform
      uint32_t bits[4] = {0u, 0u, 0u, 0u};
      if (masked) key_bits(prm.mask, bh, row, prm.Lq, prm.Lk, key0, bits);
      mbar_wait(&sm.s_full, j & 1);
      tc_fence_after();
      const float m_tile = masked ? fwd_tile_max<true>(t_row, bits) : fwd_tile_max<false>(t_row, bits);
      const float m_new = fmaxf(m_run, m_tile);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_new);
      float l_tile;
      if (prm.drop.thr != 0u) {
        const uint32_t rowkey = drop_row_key(prm.drop, bh, row);
        l_tile = masked ? fwd_tile_exp<true, true>(t_row, bits, m_safe, prm.drop, rowkey, key0) : fwd_tile_exp<false, true>(t_row, bits, m_safe, prm.drop, rowkey, key0);
      } else {
        l_tile = masked ? fwd_tile_exp<true, false>(t_row, bits, m_safe, prm.drop, 0u, key0) : fwd_tile_exp<false, false>(t_row, bits, m_safe, prm.drop, 0u, key0);
      }
      l_run = l_run * alpha + l_tile;
      // rescale the running output (previous PV MMAs are complete: s_full was committed after them).
      // tcgen05.ld/st are warp-collective (.sync.aligned): the branch must be warp-uniform
      if (j > 0 && j % kFlushTiles == 0) {
        // fold the tensor-memory accumulator into the register accumulator; the next PV MMAs restart it
        float t0[16], t1[16];
        tmem_ld16(t_row + kTmemO, t0);
        tmem_ld16(t_row + kTmemO + 16, t1);
#pragma unroll
        for (int i = 0; i < 16; ++i) o_reg[i] = (o_reg[i] + (t0[i] + t1[i])) * alpha;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o_reg[i] *= alpha;
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {          // both halves of the accumulator: P V_hi and P V_lo
            float o[16];
            tmem_ld16(t_row + kTmemO + h * 16, o);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] *= alpha;
            tmem_st16(t_row + kTmemO + h * 16, o);
          }
        }
      }
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(&sm.p_full);
    }
    // epilogue
    mbar_wait(&sm.o_full, 0);
    tc_fence_after();
    float o[16], o2[16];
    tmem_ld16(t_row + kTmemO, o);
    tmem_ld16(t_row + kTmemO + 16, o2);
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = o_reg[i] + (o[i] + o2[i]);
    if (row < prm.Lq) {
      const float inv = 1.f / l_run;   // fully masked row: 0 * inf = NaN, like softmax over all -inf
      float* dst = prm.out + ((size_t)bh * prm.Lq + row) * prm.dv;
      for (int i = 0; i < prm.dv; ++i) dst[i] = o[i] * inv;
      if (prm.lse2) prm.lse2[(size_t)bh * prm.Lq + row] = m_run + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------- attention probabilities on request
// attn[bh][i][j] = exp2(s_ij - lse2_i) (0 where masked): the reference returns this tensor
// (utils.py:325-329); it is L_q x L_k per head, so it is only produced when asked for.
__global__ void attn_probs_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ lse2,
                                  const unsigned char* __restrict__ key_mask, const unsigned char* __restrict__ full_mask, int n_mask_rows,
                                  int Lq, int Lk, int d, float qscale, DropArgs drop, float* __restrict__ attn) {
  const int bh = blockIdx.z, i = blockIdx.y;
  const uint32_t rowkey = drop_row_key(drop, bh, i);
  const float* qi = q + ((size_t)bh * Lq + i) * d;
  const float l2 = lse2[(size_t)bh * Lq + i];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Lk; j += gridDim.x * blockDim.x) {
    const float* kj = k + ((size_t)bh * Lk + j) * d;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s = fmaf(qi[c] * qscale, kj[c], s);
    bool masked = false;
    if (key_mask) masked = key_mask[(size_t)(bh % n_mask_rows) * Lk + j] != 0;
    if (!masked && full_mask) masked = full_mask[((size_t)(bh % n_mask_rows) * Lq + i) * Lk + j] != 0;
    float pr = masked ? 0.f : exp2f(s - l2);
    if (drop.thr != 0u) pr = drop_one(drop, rowkey, j) ? 0.f : pr * drop.scale;     // what multiplied V (utils.py:326-327)
    attn[((size_t)bh * Lq + i) * Lk + j] = pr;
  }
}

// ---------------------------------------------------------------------------- backward
// Per query tile (33 KB):  Q'h | Q'l | Q'^T h | Q'^T l | dOh | dOl | dO^T h | dO^T l | lse2[128] | delta[128]
// Per key tile   (24 KB):  Kh | Kl | Vh | Vl | K^T h | K^T l
// Q' = q * log2(e)/T, so S = Q' K^T is in the exp2 domain and P = exp2(S - lse2).
// With A = dL/d(q.k/T) = P (dP - delta):  dQ = A K / T,  dK = ln2 * A^T Q',  dV = P^T dO.
constexpr int kBwdQTileBytes = 8 * kOperandBytes + 1024;
constexpr int kBwdKTileBytes = 6 * kOperandBytes;

struct BwdPrepParams {
  const float* q; const float* k; const float* v; const float* dout; const float* out; const float* lse2;
  unsigned char* qtiles; unsigned char* ktiles;
  int BH, Lq, Lk, d, dv, nQt, nKt;
  float qscale;
};

__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const BwdPrepParams p) {
  const int bh = blockIdx.y, tile = blockIdx.x;
  if (tile < p.nQt) {
    unsigned char* dst = p.qtiles + ((size_t)bh * p.nQt + tile) * kBwdQTileBytes;
    const float* q = p.q + (size_t)bh * p.Lq * p.d;
    const float* go = p.dout + (size_t)bh * p.Lq * p.dv;
    write_row_tile(dst, dst + kOperandBytes, q, p.d, p.d, tile * kTileQ, p.Lq, p.qscale);
    write_col_tile(dst + 2 * kOperandBytes, dst + 3 * kOperandBytes, q, p.d, p.d, tile * kTileQ, p.Lq, p.qscale);
    write_row_tile(dst + 4 * kOperandBytes, dst + 5 * kOperandBytes, go, p.dv, p.dv, tile * kTileQ, p.Lq, 1.f);
    write_col_tile(dst + 6 * kOperandBytes, dst + 7 * kOperandBytes, go, p.dv, p.dv, tile * kTileQ, p.Lq, 1.f);
    if (threadIdx.x < kTileQ) {
      const int row = tile * kTileQ + threadIdx.x;
      float l2 = 0.f, dl = 0.f;
      if (row < p.Lq) {
        l2 = p.lse2[(size_t)bh * p.Lq + row];
        const float* o = p.out + ((size_t)bh * p.Lq + row) * p.dv;
        for (int c = 0; c < p.dv; ++c) dl = fmaf(go[(size_t)row * p.dv + c], o[c], dl);
      }
      float* st = reinterpret_cast<float*>(dst + 8 * kOperandBytes);
      st[threadIdx.x] = l2;
      st[kTileQ + threadIdx.x] = dl;
    }
  }
  if (tile < p.nKt) {
    unsigned char* dst = p.ktiles + ((size_t)bh * p.nKt + tile) * kBwdKTileBytes;
    const float* k = p.k + (size_t)bh * p.Lk * p.d;
    const float* v = p.v + (size_t)bh * p.Lk * p.dv;
    write_row_tile(dst, dst + kOperandBytes, k, p.d, p.d, tile * kTileK, p.Lk, 1.f);
    write_row_tile(dst + 2 * kOperandBytes, dst + 3 * kOperandBytes, v, p.dv, p.dv, tile * kTileK, p.Lk, 1.f);
    write_col_tile(dst + 4 * kOperandBytes, dst + 5 * kOperandBytes, k, p.d, p.d, tile * kTileK, p.Lk, 1.f);
  }
}

struct BwdParams {
  const unsigned char* qtiles; const unsigned char* ktiles;
  MaskArgs mask;
  DropArgs drop;
  float* dq; float* dk; float* dv;
  int BH, Lq, Lk, d, dv_dim, nQt, nKt;
  float inv_t;
};

// ---- pipelined backward kernel (one template, two roles) -----------------------------------------------------------
// The CTA owns one 128-row tile (DKDV = false: queries, produces dQ; DKDV = true: keys, produces dK and dV) and streams
// the 128-wide tiles of the other side, each processed as two 64-column sub-tiles `u`.  TMEM stage u&1 (128 columns):
//   tensor pipe   S(u) -> cols [0,64), dP(u) -> cols [64,128)                     (6 MMAs, N = 64, operands in smem)
//   math warps    tcgen05.ld S, dP -> P = exp2(S - lse2), dA = P (dP - delta) -> bf16 hi/lo pairs written back with
//                 tcgen05.st OVER the S/dP columns the same thread has just read (P over S, dA over dP)
//   tensor pipe   acc += P/dA(u) x B  with the A operand read from TMEM (12 / 24 MMAs, N = 16), then S, dP(u+2)
// so the exponentials of sub-tile u+1 run while the tensor pipe works on u and u+2, and P/dA never touch shared memory
// (the shared-memory version spent more than half of the SM's smem bandwidth on 16 KB operand re-reads by N = 16 MMAs).
// Tiles without masked entries take a branch-free path (padded queries / keys contribute exact zeros through their
// zero operand rows).
constexpr int kSub = 64;
constexpr int kBwdStages = 3;
constexpr int kTmStage = 128;                           // TMEM columns per stage: S | dP
constexpr int kTmAcc = 256;                             // accumulators: [which (dV|dQ, dK)][chain 0/1][32 columns = x B_hi | x B_lo]

struct BwdSmem {
  unsigned char own[4 * kOperandBytes];                 // dQ: Q'h | Q'l | dOh | dOl      dK/dV: Kh | Kl | Vh | Vl
  unsigned char oth[kBwdStages][kBwdQTileBytes];        // streamed tiles of the other side (24 KB key / 33 KB query tiles)
  uint64_t c_full, o_full;
  uint64_t s_full[2], a_full[2];
  uint64_t t_full[kBwdStages], t_empty[kBwdStages];
  uint32_t tmem_base;
};

struct BwdMaskCtx {           // per-thread mask state of the slow path
  uint32_t bits;              // dQ: masked keys among this thread's 32 columns
  bool row_masked;            // dK/dV: this key is masked for every query
  const unsigned char* fm;    // dK/dV: full mask base for this head (or null)
  int q0, Lq, Lk, own_row;
  // dropout: dQ role: hash key of the thread's query row and the first key of its 32 columns; dK/dV role: bh (queries vary)
  uint32_t rowkey; int k0; int bh;
};

// one thread, 32 columns: registers of S and dP -> packed bf16 hi/lo pairs of P and dA
template <bool DKDV, bool MASKED, bool DROP>
__device__ __forceinline__ void bwd_math(const uint32_t (&sr)[32], const uint32_t (&dr)[32], const float* __restrict__ st_l2, const float* __restrict__ st_dl,
                                         float l2_row, float dl_row, const BwdMaskCtx& mc, const DropArgs& drop, uint32_t (&ph)[16], uint32_t (&pl)[16],
                                         uint32_t (&ah)[16], uint32_t (&al)[16]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float l2v[4] = {l2_row, l2_row, l2_row, l2_row}, dlv[4] = {dl_row, dl_row, dl_row, dl_row};
    if (DKDV) {
      const float4 a = *reinterpret_cast<const float4*>(st_l2 + g * 4);
      const float4 d = *reinterpret_cast<const float4*>(st_dl + g * 4);
      l2v[0] = a.x; l2v[1] = a.y; l2v[2] = a.z; l2v[3] = a.w;
      dlv[0] = d.x; dlv[1] = d.y; dlv[2] = d.z; dlv[3] = d.w;
    }
    float p4[4], a4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = g * 4 + e;
      float p = ex2_approx(__uint_as_float(sr[i]) - l2v[e]);
      if (MASKED) {
        bool mk;
        if (DKDV) {
          const int qi = mc.q0 + i;
          mk = mc.row_masked || qi >= mc.Lq;
          if (mc.fm && !mk) mk = mc.fm[(size_t)qi * mc.Lk + mc.own_row] != 0;
        } else {
          mk = (mc.bits >> i) & 1u;
        }
        if (mk) p = 0.f;
      }
      // dropout: the probability that multiplied V in the forward is p * keep / (1 - p_drop); its gradient dP gets the
      // same factor, delta = rowsum(dO * O) is unchanged
      float dp = __uint_as_float(dr[i]);
      float pk = p;
      if (DROP) {
        bool dropped;
        if (DKDV) dropped = drop_one(drop, drop_row_key(drop, mc.bh, mc.q0 + i), mc.own_row);
        else dropped = drop_one(drop, mc.rowkey, mc.k0 + i);
        const float f = dropped ? 0.f : drop.scale;
        pk = p * f;
        dp *= f;
      }
      p4[e] = pk;
      a4[e] = p * (dp - dlv[e]);
    }
    if (DKDV) {
      split2(p4[0], p4[1], ph[g * 2], pl[g * 2]);
      split2(p4[2], p4[3], ph[g * 2 + 1], pl[g * 2 + 1]);
    }
    split2(a4[0], a4[1], ah[g * 2], al[g * 2]);
    split2(a4[2], a4[3], ah[g * 2 + 1], al[g * 2 + 1]);
  }
}

template <bool DKDV>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_kernel(const BwdParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int own_t = blockIdx.x, bh = blockIdx.y;
  const int nT = DKDV ? prm.nQt : prm.nKt;            // streamed tiles
  const int U = 2 * nT;                               // sub-tiles
  constexpr uint32_t kOthBytes = DKDV ? kBwdQTileBytes : kBwdKTileBytes;
  const unsigned char* own_src = DKDV ? prm.ktiles + ((size_t)bh * prm.nKt + own_t) * kBwdKTileBytes
                                      : prm.qtiles + ((size_t)bh * prm.nQt + own_t) * kBwdQTileBytes;
  const unsigned char* oth_src = DKDV ? prm.qtiles + (size_t)bh * prm.nQt * kBwdQTileBytes : prm.ktiles + (size_t)bh * prm.nKt * kBwdKTileBytes;

  if (threadIdx.x == 0) {
    mbar_init(&sm.c_full, 1); mbar_init(&sm.o_full, 2);
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.s_full[s], 1); mbar_init(&sm.a_full[s], 256); }
    for (int s = 0; s < kBwdStages; ++s) { mbar_init(&sm.t_full[s], 1); mbar_init(&sm.t_empty[s], 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sm.tmem_base, kBwdTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&sm.c_full, 4 * kOperandBytes);
      if (DKDV) {
        tma_bulk_g2s(sm.own, own_src, 4 * kOperandBytes, &sm.c_full);                              // Kh | Kl | Vh | Vl
      } else {
        tma_bulk_g2s(sm.own, own_src, 2 * kOperandBytes, &sm.c_full);                              // Q'h | Q'l
        tma_bulk_g2s(sm.own + 2 * kOperandBytes, own_src + 4 * kOperandBytes, 2 * kOperandBytes, &sm.c_full);   // dOh | dOl
      }
      for (int i = 0; i < nT; ++i) {
        const int st = i % kBwdStages;
        mbar_wait(&sm.t_empty[st], ((i / kBwdStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&sm.t_full[st], kOthBytes);
        tma_bulk_g2s(sm.oth[st], oth_src + (size_t)i * kOthBytes, kOthBytes, &sm.t_full[st]);
      }
    }
  } else if (warp <= 2) {
    // ===================== MMA issuers: chain c = warp - 1 owns the sub-tiles u with u % 2 == c, TMEM stage c and its own
    // accumulators (summed in the epilogue), so the two chains never wait for each other and the per-MMA issue cost
    // (descriptor arithmetic + R2UR moves of a single thread) is split between two warps
    {
      const int c = warp - 1;
      const uint32_t el = elect_one();
      constexpr uint32_t idesc_s = make_idesc(128, kSub);
      constexpr uint32_t idesc_o2 = make_idesc(128, 2 * kDP);    // A x [B_hi ; B_lo] -> 32 columns
      constexpr uint32_t idesc_o1 = make_idesc(128, kDP);
      // descriptors differ from their base only in the 16-byte address field: desc(addr + x) = desc(addr) + (x >> 4)
      const uint64_t dA0h = make_desc(smem_u32(sm.own), 128, kSboQK);                               // A of S:  Q' (dQ) / K (dK,dV)
      const uint64_t dA0l = dA0h + (kOperandBytes >> 4);
      const uint64_t dA1h = dA0h + (2 * kOperandBytes >> 4), dA1l = dA0h + (3 * kOperandBytes >> 4);   // A of dP: dO (dQ) / V (dK,dV)
      const uint64_t dRow0 = make_desc(smem_u32(sm.oth[0]), 128, kSboQK);                           // row operands of a streamed tile (B of S / dP)
      const uint64_t dCol0 = make_desc(smem_u32(sm.oth[0]), 128, kSboP);                            // transposed operands (B of the accumulations)
      constexpr uint32_t oP = (DKDV ? 4 * kOperandBytes : 2 * kOperandBytes) >> 4;
      constexpr uint32_t oB0 = (DKDV ? 6 * kOperandBytes : 4 * kOperandBytes) >> 4;                 // dO^T (dV) / K^T (dQ)
      constexpr uint32_t oB1 = (2 * kOperandBytes) >> 4;                                            // Q'^T (dK)
      constexpr uint32_t oLo = kOperandBytes >> 4;
      constexpr uint32_t kStageStride = kBwdQTileBytes >> 4;
      const uint32_t tS = tmem + c * kTmStage, tP = tS + kSub;
      const uint32_t acc0 = tmem + kTmAcc + c * 32, acc1 = tmem + kTmAcc + 64 + c * 32;
      auto scores = [&](int n) {                          // S, dP of sub-tile c of streamed tile n -> TMEM stage c
        const int st = n % kBwdStages;
        mbar_wait(&sm.t_full[st], (n / kBwdStages) & 1);
        tc_fence_after();
        const uint64_t ob = dRow0 + (uint32_t)st * kStageStride + (uint32_t)c * (2048u >> 4);       // 64 rows = 8 row groups of 256 B
        umma_bf16_e(el, tS, dA0h, ob, idesc_s, 0);
        umma_bf16_e(el, tS, dA0h, ob + oLo, idesc_s, 1);
        umma_bf16_e(el, tS, dA0l, ob, idesc_s, 1);
        umma_bf16_e(el, tP, dA1h, ob + oP, idesc_s, 0);
        umma_bf16_e(el, tP, dA1h, ob + oP + oLo, idesc_s, 1);
        umma_bf16_e(el, tP, dA1l, ob + oP, idesc_s, 1);
        umma_commit_e(el, &sm.s_full[c]);
      };
      mbar_wait(&sm.c_full, 0);
      tc_fence_after();
      scores(0);
      for (int n = 0; n < nT; ++n) {
        const int st = n % kBwdStages;
        const uint64_t ob = dCol0 + (uint32_t)st * kStageStride + (uint32_t)c * (1024u >> 4);       // 64 k = 8 core matrices of 128 B
        mbar_wait(&sm.a_full[c], n & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kSub / 16; ++kk) {
          // packed operand columns of k-step kk inside the stage: math warp `half` = kk / 2 owns columns [half*32, +32) of
          // the S part (P hi | P lo, 16 columns each) and of the dP part (dA hi | dA lo)
          const uint32_t ca = (kk >> 1) * 32 + (kk & 1) * 8;
          const uint32_t ko = (kk * 256) >> 4;
          const uint32_t acc = (n > 0 || kk > 0) ? 1u : 0u;
          if (DKDV) {
            umma_bf16_ts_e(el, acc0, tS + ca, ob + oB0 + ko, idesc_o2, acc);          // P_hi x [dO_hi ; dO_lo]
            umma_bf16_ts_e(el, acc0, tS + ca + 16, ob + oB0 + ko, idesc_o1, 1);       // P_lo x dO_hi
            umma_bf16_ts_e(el, acc1, tP + ca, ob + oB1 + ko, idesc_o2, acc);          // dA_hi x [Q'_hi ; Q'_lo]
            umma_bf16_ts_e(el, acc1, tP + ca + 16, ob + oB1 + ko, idesc_o1, 1);       // dA_lo x Q'_hi
          } else {
            umma_bf16_ts_e(el, acc0, tP + ca, ob + oB0 + ko, idesc_o2, acc);          // dA_hi x [K_hi ; K_lo]
            umma_bf16_ts_e(el, acc0, tP + ca + 16, ob + oB0 + ko, idesc_o1, 1);       // dA_lo x K_hi
          }
        }
        umma_commit_e(el, &sm.t_empty[st]);                     // this chain is done with the streamed tile (the barrier counts both chains)
        if (n + 1 < nT) scores(n + 1);                    // in issue order behind the MMAs that read this TMEM stage
      }
      umma_commit_e(el, &sm.o_full);
    }
  } else if (warp >= 4) {
    // ===================== 8 math warps: TMEM quarter = warp % 4, 32-column half = (warp - 4) / 4 =====================
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;                  // row inside the own tile
    const int own_row = own_t * 128 + r;
    const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);
    float l2_row = 0.f, dl_row = 0.f;
    BwdMaskCtx mc;
    mc.bits = 0u; mc.row_masked = false; mc.fm = nullptr; mc.q0 = 0; mc.Lq = prm.Lq; mc.Lk = prm.Lk; mc.own_row = own_row;
    mc.bh = bh; mc.k0 = 0; mc.rowkey = DKDV ? 0u : drop_row_key(prm.drop, bh, own_row);
    bool masked = false;
    if (DKDV) {
      mc.row_masked = own_row >= prm.Lk;
      if (!mc.row_masked && prm.mask.key_mask) mc.row_masked = prm.mask.key_mask[(size_t)(bh % prm.mask.n_mask_rows) * prm.Lk + own_row] != 0;
      mc.fm = prm.mask.full_mask ? prm.mask.full_mask + (size_t)(bh % prm.mask.n_mask_rows) * prm.Lq * prm.Lk : nullptr;
      // padded rows (own_row >= Lk) may stay on the fast path: their outputs are never stored
      masked = __any_sync(0xffffffffu, mc.row_masked && own_row < prm.Lk) || mc.fm != nullptr;
    } else {
      const float* stats = reinterpret_cast<const float*>(own_src + 8 * kOperandBytes);
      l2_row = __ldg(stats + r);
      dl_row = __ldg(stats + kTileQ + r);
    }
    uint32_t bits[4] = {0u, 0u, 0u, 0u};
    for (int u = 0; u < U; ++u) {
      const int b = u & 1, n = u >> 1, st = n % kBwdStages;
      const int col0 = b * kSub + half * 32;             // first column (inside the streamed 128-tile) of this thread's 32
      if (!DKDV && b == 0) {
        masked = tile_needs_mask(prm.mask, bh, prm.Lk, n * kTileK, lane);
        if (masked) key_bits(prm.mask, bh, own_row, prm.Lq, prm.Lk, n * kTileK, bits);
      }
      if (DKDV && b == 0) mbar_wait(&sm.t_full[st], (n / kBwdStages) & 1);   // lse2 / delta of this query tile are in smem
      mbar_wait(&sm.s_full[b], n & 1);
      tc_fence_after();
      const uint32_t t_s = t_row + b * kTmStage + half * 32, t_p = t_s + kSub;
      uint32_t sr[32], dr[32];
      tmem_ld32_issue(t_s, sr);
      tmem_ld32_issue(t_p, dr);
      tmem_ld32_wait(sr);
      tmem_ld32_wait(dr);
      const float* stats = reinterpret_cast<const float*>(sm.oth[st] + 8 * kOperandBytes);
      uint32_t ph[16], pl[16], ah[16], al[16];
      mc.q0 = n * kTileQ + col0;
      mc.k0 = n * kTileK + col0;
      const bool dropping = prm.drop.thr != 0u;
      if (!masked) {
        if (dropping) bwd_math<DKDV, false, true>(sr, dr, stats + col0, stats + kTileQ + col0, l2_row, dl_row, mc, prm.drop, ph, pl, ah, al);
        else bwd_math<DKDV, false, false>(sr, dr, stats + col0, stats + kTileQ + col0, l2_row, dl_row, mc, prm.drop, ph, pl, ah, al);
      } else {
        mc.bits = b ? (half ? bits[3] : bits[2]) : (half ? bits[1] : bits[0]);
        if (dropping) bwd_math<DKDV, true, true>(sr, dr, stats + col0, stats + kTileQ + col0, l2_row, dl_row, mc, prm.drop, ph, pl, ah, al);
        else bwd_math<DKDV, true, false>(sr, dr, stats + col0, stats + kTileQ + col0, l2_row, dl_row, mc, prm.drop, ph, pl, ah, al);
      }
      if (DKDV) {
        tmem_st16_issue(t_s, ph);
        tmem_st16_issue(t_s + 16, pl);
      }
      tmem_st16_issue(t_p, ah);
      tmem_st16_issue(t_p + 16, al);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&sm.a_full[b]);
    }
    mbar_wait(&sm.o_full, 0);
    tc_fence_after();
    if (DKDV || half == 0) {
      // result = sum over the two chains and the [x B_hi | x B_lo] column halves of this accumulator
      float o[16], t[16];
      const uint32_t ta = t_row + kTmAcc + (DKDV && half == 1 ? 64 : 0);
      tmem_ld16(ta, o);
#pragma unroll
      for (int part = 1; part < 4; ++part) {
        tmem_ld16(ta + part * 16, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] += t[i];
      }
      if (!DKDV) {
        if (own_row < prm.Lq) {
          float* dst = prm.dq + ((size_t)bh * prm.Lq + own_row) * prm.d;
          for (int i = 0; i < prm.d; ++i) dst[i] = o[i] * prm.inv_t;
        }
      } else if (own_row < prm.Lk) {
        if (half == 0) {
          float* dst = prm.dv + ((size_t)bh * prm.Lk + own_row) * prm.dv_dim;
          for (int c = 0; c < prm.dv_dim; ++c) dst[c] = o[c];
        } else {
          float* dst = prm.dk + ((size_t)bh * prm.Lk + own_row) * prm.d;
          for (int c = 0; c < prm.d; ++c) dst[c] = o[c] * 0.6931471805599453f;   // ln 2: S was in the exp2 domain
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kBwdTmemCols);
}

// ---------------------------------------------------------------------------- TMEM read-rate probe (diagnostics)
template <int X>
__global__ void __launch_bounds__(512, 1) tmem_ld_probe_kernel(int iters, uint32_t* sink) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&s_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_row = s_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * X + (warp >> 2) * X) & 255);
    if (X == 32) {
      uint32_t r[32];
      tmem_ld32_issue(t_row + col, r);
      tmem_ld32_wait(r);
      acc ^= r[0] ^ r[31];
    } else {
      float v[16];
      tmem_ld16(t_row + col, v);
      acc ^= __float_as_uint(v[0]) ^ __float_as_uint(v[15]);
    }
  }
  if (acc == 0x12345u) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_base, 512);
}

int check_attn(int BH, int Lq, int Lk, int d, int dv) {
  ISA_CHECK_ARG(BH > 0 && Lq > 0 && Lk > 0, "attention: non-positive size (BH=%d Lq=%d Lk=%d)", BH, Lq, Lk);
  ISA_CHECK_ARG(d > 0 && d <= kDP && dv > 0 && dv <= kDP, "attention: head widths must be <= %d (d_k=%d, d_v=%d)", kDP, d, dv);
  ISA_CHECK_ARG(BH <= 65535, "attention: n_head*batch=%d > 65535", BH);
  return ISA_OK;
}

// p in [0, 1): quantised to 1/65536 (the keep probability used for the 1/(1-p) scale is the quantised one, so the
// estimator stays unbiased); seed: any 64-bit value, folded to 32 bits
int make_drop(float p, unsigned long long seed, DropArgs* d) {
  ISA_CHECK_ARG(p >= 0.f && p < 1.f, "attention: dropout probability %f not in [0, 1)", (double)p);
  uint32_t thr = (uint32_t)((double)p * 65536.0 + 0.5);
  if (thr > 65535u) thr = 65535u;
  d->thr = thr;
  d->scale = (float)(65536.0 / (65536.0 - (double)thr));
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull;       // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  d->seed = (uint32_t)z ^ (uint32_t)(z >> 32);
  return ISA_OK;
}

size_t fwd_ws_bytes(int BH, int Lq, int Lk) {
  const size_t nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  return isa_align_up((size_t)BH * nQt * kQTileBytes, 1024) + isa_align_up((size_t)BH * nKt * kKvTileBytes, 1024);
}
size_t bwd_ws_bytes(int BH, int Lq, int Lk) {
  const size_t nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  return isa_align_up((size_t)BH * nQt * kBwdQTileBytes, 1024) + isa_align_up((size_t)BH * nKt * kBwdKTileBytes, 1024);
}

}  // namespace

extern "C" {

size_t isa_attention_workspace_bytes(int BH, int Lq, int Lk) {
  if (BH <= 0 || Lq <= 0 || Lk <= 0) return 0;
  const size_t a = fwd_ws_bytes(BH, Lq, Lk), b = bwd_ws_bytes(BH, Lq, Lk);
  return a > b ? a : b;
}

// q [BH][Lq][d], k [BH][Lk][d], v [BH][Lk][dv] fp32; temperature T (= sqrt(d_k) in the reference);
// key_mask [n_mask_rows][Lk] u8 (1 = masked, broadcast over queries; row = bh % n_mask_rows) and/or
// full_mask [n_mask_rows][Lq][Lk] u8; out [BH][Lq][dv]; lse2 [BH][Lq] (may be NULL).
int isa_attention_fwd(const float* q, const float* k, const float* v, int BH, int Lq, int Lk, int d, int dv, float temperature,
                      const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows,
                      float dropout_p, unsigned long long dropout_seed,
                      float* out, float* lse2, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, dv);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && v && out && workspace, "attention_fwd: null pointer");
  DropArgs drop;
  rc = make_drop(dropout_p, dropout_seed, &drop);
  if (rc) return rc;
  ISA_CHECK_ARG(temperature > 0.f, "attention_fwd: temperature must be positive");
  ISA_CHECK_ARG(!(key_mask || full_mask) || n_mask_rows > 0, "attention_fwd: n_mask_rows must be positive with a mask");
  if (workspace_bytes < isa_attention_workspace_bytes(BH, Lq, Lk)) {
    isa_set_error("attention_fwd: workspace too small");
    return ISA_ERR_WORKSPACE;
  }
  IsaDeviceInfo di;
  rc = isa_device_info(&di);
  if (rc) return rc;
  const int nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  unsigned char* qblk = (unsigned char*)workspace;
  unsigned char* kvblk = qblk + isa_align_up((size_t)BH * nQt * kQTileBytes, 1024);
  PrepParams pp;
  pp.q = q; pp.k = k; pp.v = v; pp.qblk = qblk; pp.kvblk = kvblk;
  pp.BH = BH; pp.Lq = Lq; pp.Lk = Lk; pp.d = d; pp.dv = dv; pp.nQt = nQt; pp.nKt = nKt;
  pp.qscale = 1.4426950408889634f / temperature;
  attn_prep_kernel<<<dim3(nQt > nKt ? nQt : nKt, BH), 256, 0, stream>>>(pp);
  ISA_CUDA(cudaGetLastError());
  FwdParams fp;
  fp.qblk = qblk; fp.kvblk = kvblk;
  fp.mask.key_mask = key_mask; fp.mask.full_mask = full_mask; fp.mask.n_mask_rows = n_mask_rows > 0 ? n_mask_rows : 1;
  fp.drop = drop;
  fp.out = out; fp.lse2 = lse2; fp.BH = BH; fp.Lq = Lq; fp.Lk = Lk; fp.dv = dv; fp.nQt = nQt; fp.nKt = nKt;
  const size_t smem = sizeof(FwdSmem) + 1024;
  ISA_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<<<dim3(nQt, BH), kFwdThreads, smem, stream>>>(fp);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// attn [BH][Lq][Lk] = softmax probabilities (0 where masked) from the saved lse2.
int isa_attention_probs(const float* q, const float* k, const float* lse2, int BH, int Lq, int Lk, int d, float temperature,
                        const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows,
                        float dropout_p, unsigned long long dropout_seed, float* attn, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, 1);
  if (rc) return rc;
  DropArgs drop;
  rc = make_drop(dropout_p, dropout_seed, &drop);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && lse2 && attn, "attention_probs: null pointer");
  ISA_CHECK_ARG(Lq <= 65535, "attention_probs: Lq=%d > 65535", Lq);
  int gx = (Lk + 255) / 256;
  if (gx > 64) gx = 64;
  attn_probs_kernel<<<dim3(gx, Lq, BH), 256, 0, stream>>>(q, k, lse2, key_mask, full_mask, n_mask_rows > 0 ? n_mask_rows : 1, Lq, Lk, d,
                                                         1.4426950408889634f / temperature, drop, attn);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// dq/dk/dv from (q,k,v,out,dout,lse2); workspace: isa_attention_workspace_bytes (contents of forward are not needed).
int isa_attention_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout, const float* lse2,
                      int BH, int Lq, int Lk, int d, int dv, float temperature,
                      const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows,
                      float dropout_p, unsigned long long dropout_seed,
                      float* dq, float* dk, float* dvv, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, dv);
  if (rc) return rc;
  DropArgs drop;
  rc = make_drop(dropout_p, dropout_seed, &drop);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && v && out && dout && lse2 && dq && dk && dvv && workspace, "attention_bwd: null pointer");
  if (workspace_bytes < isa_attention_workspace_bytes(BH, Lq, Lk)) {
    isa_set_error("attention_bwd: workspace too small");
    return ISA_ERR_WORKSPACE;
  }
  const int nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  unsigned char* qtiles = (unsigned char*)workspace;
  unsigned char* ktiles = qtiles + isa_align_up((size_t)BH * nQt * kBwdQTileBytes, 1024);
  BwdPrepParams pp;
  pp.q = q; pp.k = k; pp.v = v; pp.dout = dout; pp.out = out; pp.lse2 = lse2; pp.qtiles = qtiles; pp.ktiles = ktiles;
  pp.BH = BH; pp.Lq = Lq; pp.Lk = Lk; pp.d = d; pp.dv = dv; pp.nQt = nQt; pp.nKt = nKt;
  pp.qscale = 1.4426950408889634f / temperature;
  attn_bwd_prep_kernel<<<dim3(nQt > nKt ? nQt : nKt, BH), 256, 0, stream>>>(pp);
  ISA_CUDA(cudaGetLastError());
  BwdParams bp;
  bp.qtiles = qtiles; bp.ktiles = ktiles;
  bp.mask.key_mask = key_mask; bp.mask.full_mask = full_mask; bp.mask.n_mask_rows = n_mask_rows > 0 ? n_mask_rows : 1;
  bp.drop = drop;
  bp.dq = dq; bp.dk = dk; bp.dv = dvv; bp.BH = BH; bp.Lq = Lq; bp.Lk = Lk; bp.d = d; bp.dv_dim = dv; bp.nQt = nQt; bp.nKt = nKt;
  bp.inv_t = 1.f / temperature;
  const size_t smem = sizeof(BwdSmem) + 1024;
  ISA_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ISA_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_kernel<false><<<dim3(nQt, BH), kBwdThreads, smem, stream>>>(bp);   // dQ
  attn_bwd_kernel<true><<<dim3(nKt, BH), kBwdThreads, smem, stream>>>(bp);    // dK, dV
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// Diagnostics: bytes per clock per SM sustained by `warps` (4, 8 or 16) resident warps each issuing back-to-back
// tcgen05.ld.32x32b.x{cols} (cols = 16 or 32) + wait::ld -- the TMEM read rate the softmax / backward math warps
// are sized against.  Synchronises the device; scratch = 4 bytes of device memory.
int isa_selftest_tmem_ld_rate(int warps, int cols, float sm_clock_mhz, void* scratch, float* h_bytes_per_clk_per_sm) {
  ISA_CHECK_ARG((warps == 4 || warps == 8 || warps == 16) && (cols == 16 || cols == 32) && sm_clock_mhz > 0 && scratch && h_bytes_per_clk_per_sm,
                "selftest_tmem_ld_rate: bad argument");
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  const int iters = 1 << 14;
  cudaEvent_t e0, e1;
  ISA_CUDA(cudaEventCreate(&e0));
  ISA_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    ISA_CUDA(cudaEventRecord(e0, 0));
    if (cols == 32) tmem_ld_probe_kernel<32><<<di.num_sms, warps * 32>>>(iters, reinterpret_cast<uint32_t*>(scratch));
    else tmem_ld_probe_kernel<16><<<di.num_sms, warps * 32>>>(iters, reinterpret_cast<uint32_t*>(scratch));
    ISA_CUDA(cudaEventRecord(e1, 0));
    ISA_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ISA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double bytes_per_sm = (double)iters * warps * 32.0 * cols * 4.0;
  *h_bytes_per_clk_per_sm = (float)(bytes_per_sm / ((double)best * 1e-3 * sm_clock_mhz * 1e6));
  return ISA_OK;
}

}  // extern "C"
