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
// Backward: recompute-based kernel on the CUDA cores (fp32) -- dQ, dK, dV from (q,k,v,dO,lse).
#include "isa_common.cuh"
#include <cuda_bf16.h>
#include <math.h>

namespace {

constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr int kDP = 16;            // padded head width
constexpr int kStages = 2;
constexpr int kOperandBytes = kTileQ * kDP * 2;     // 4096: one bf16 [128][16] operand tile
constexpr int kKvTileBytes = 4 * kOperandBytes;     // Kh | Kl | Vh^T | Vl^T
constexpr int kQTileBytes = 2 * kOperandBytes;      // Qh | Ql
constexpr int kPBytes = kTileQ * kTileK * 2;        // 32768: one bf16 [128][128] operand
constexpr int kFwdThreads = 192;
constexpr int kTmemCols = 256;                      // S: [0,128), O: [128,144)
constexpr int kTmemO = 128;

// canonical K-major no-swizzle layouts (units: bytes).  Element (row r, k):
//   (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2     (core matrix = 8 rows x 16 B, contiguous)
__host__ __device__ inline int kmajor_off(int r, int k, int sbo) { return (r >> 3) * sbo + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }
constexpr int kSboQK = 256;    // [128 rows][16 k]: 2 core matrices along K
constexpr int kSboP = 2048;    // [128 rows][128 k]: 16 core matrices along K (also V^T: [16 rows][128 k])

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);  // version = 1 (Blackwell), layout_type = SWIZZLE_NONE, base_offset = 0
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---------------------------------------------------------------------------- prep
// q [BH][Lq][d], k [BH][Lk][d], v [BH][Lk][dv] fp32  ->  tile-blocked bf16 hi/lo operands
struct PrepParams {
  const float* q; const float* k; const float* v;
  unsigned char* qblk;   // [BH][nQt][Qh 4K | Ql 4K]
  unsigned char* kvblk;  // [BH][nKt][Kh | Kl | Vh^T | Vl^T]
  int BH, Lq, Lk, d, dv, nQt, nKt;
  float qscale;          // log2(e) / temperature
};

__global__ void attn_prep_kernel(const PrepParams p) {
  const int bh = blockIdx.y;
  const int tile = blockIdx.x;
  // one thread per (row, 8-element k group): 128 rows x 2 groups = 256 threads
  const int r = threadIdx.x >> 1, kg = threadIdx.x & 1;
  if (tile < p.nQt) {
    unsigned char* dst = p.qblk + ((size_t)bh * p.nQt + tile) * kQTileBytes;
    const int row = tile * kTileQ + r;
    uint32_t hi4[4], lo4[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float x0 = 0.f, x1 = 0.f;
      const int c0 = kg * 8 + e;
      if (row < p.Lq) {
        if (c0 < p.d) x0 = p.q[((size_t)bh * p.Lq + row) * p.d + c0] * p.qscale;
        if (c0 + 1 < p.d) x1 = p.q[((size_t)bh * p.Lq + row) * p.d + c0 + 1] * p.qscale;
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
      hi4[e >> 1] = pack2(h0, h1); lo4[e >> 1] = pack2(l0, l1);
    }
    const int off = kmajor_off(r, kg * 8, kSboQK);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    *reinterpret_cast<uint4*>(dst + kOperandBytes + off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
  }
  if (tile < p.nKt) {
    unsigned char* dst = p.kvblk + ((size_t)bh * p.nKt + tile) * kKvTileBytes;
    const int row = tile * kTileK + r;
    uint32_t hi4[4], lo4[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float x0 = 0.f, x1 = 0.f;
      const int c0 = kg * 8 + e;
      if (row < p.Lk) {
        if (c0 < p.d) x0 = p.k[((size_t)bh * p.Lk + row) * p.d + c0];
        if (c0 + 1 < p.d) x1 = p.k[((size_t)bh * p.Lk + row) * p.d + c0 + 1];
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
      hi4[e >> 1] = pack2(h0, h1); lo4[e >> 1] = pack2(l0, l1);
    }
    const int off = kmajor_off(r, kg * 8, kSboQK);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    *reinterpret_cast<uint4*>(dst + kOperandBytes + off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
    // V^T: operand rows = value channel n (16), k = key within the tile (128).
    // thread -> (n = threadIdx.x / 16, key group kgv = threadIdx.x % 16): 8 consecutive keys
    const int n = threadIdx.x >> 4, kgv = threadIdx.x & 15;
    uint32_t vh[4], vl[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float x0 = 0.f, x1 = 0.f;
      const int key0 = tile * kTileK + kgv * 8 + e;
      if (n < p.dv) {
        if (key0 < p.Lk) x0 = p.v[((size_t)bh * p.Lk + key0) * p.dv + n];
        if (key0 + 1 < p.Lk) x1 = p.v[((size_t)bh * p.Lk + key0 + 1) * p.dv + n];
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(x0, h0, l0); split_bf16(x1, h1, l1);
      vh[e >> 1] = pack2(h0, h1); vl[e >> 1] = pack2(l0, l1);
    }
    const int offv = kmajor_off(n, kgv * 8, kSboP);
    *reinterpret_cast<uint4*>(dst + 2 * kOperandBytes + offv) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
    *reinterpret_cast<uint4*>(dst + 3 * kOperandBytes + offv) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
  }
}

// ---------------------------------------------------------------------------- forward
struct FwdParams {
  const unsigned char* qblk;
  const unsigned char* kvblk;
  const unsigned char* key_mask;   // [n_mask_rows][Lk] 1 = masked, or null (row = bh % n_mask_rows)
  const unsigned char* full_mask;  // [n_mask_rows][Lq][Lk] or null
  int n_mask_rows;
  float* out;    // [BH][Lq][dv]
  float* lse2;   // [BH][Lq]  log2-domain log-sum-exp of the scaled scores (for backward)
  int BH, Lq, Lk, dv, nQt, nKt;
};

struct FwdSmem {
  unsigned char q[kQTileBytes];
  unsigned char kv[kStages][kKvTileBytes];
  unsigned char p_hi[kPBytes];
  unsigned char p_lo[kPBytes];
  uint64_t q_full, s_full, p_full, o_full;
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_kernel(const FwdParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int nKt = prm.nKt;

  if (threadIdx.x == 0) {
    mbar_init(&sm.q_full, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, 128);
    mbar_init(&sm.o_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.kv_full[s], 1); mbar_init(&sm.kv_empty[s], 1); }
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
        const int st = j % kStages;
        mbar_wait(&sm.kv_empty[st], ((j / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&sm.kv_full[st], kKvTileBytes);
        tma_bulk_g2s(sm.kv[st], prm.kvblk + ((size_t)bh * nKt + j) * kKvTileBytes, kKvTileBytes, &sm.kv_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kTileQ, kTileK);
      constexpr uint32_t idesc_o = make_idesc(kTileQ, kDP);
      const uint32_t q_hi = smem_u32(sm.q), q_lo = q_hi + kOperandBytes;
      const uint32_t p_hi = smem_u32(sm.p_hi), p_lo = smem_u32(sm.p_lo);
      mbar_wait(&sm.q_full, 0);
      for (int j = 0; j < nKt; ++j) {
        const int st = j % kStages;
        const uint32_t kvb = smem_u32(sm.kv[st]);
        mbar_wait(&sm.kv_full[st], (j / kStages) & 1);
        tc_fence_after();
        // S = Qh Kh^T + Qh Kl^T + Ql Kh^T
        umma_bf16(tmem, make_desc(q_hi, 128, kSboQK), make_desc(kvb, 128, kSboQK), idesc_s, 0);
        umma_bf16(tmem, make_desc(q_hi, 128, kSboQK), make_desc(kvb + kOperandBytes, 128, kSboQK), idesc_s, 1);
        umma_bf16(tmem, make_desc(q_lo, 128, kSboQK), make_desc(kvb, 128, kSboQK), idesc_s, 1);
        umma_commit(&sm.s_full);
        // O += Ph Vh + Ph Vl + Pl Vh   (P written by the softmax warps)
        mbar_wait(&sm.p_full, j & 1);
        tc_fence_after();
        const uint32_t vh = kvb + 2 * kOperandBytes, vl = kvb + 3 * kOperandBytes;
#pragma unroll
        for (int kk = 0; kk < kTileK / 16; ++kk) {
          const uint32_t ko = kk * 256;
          umma_bf16(tmem + kTmemO, make_desc(p_hi + ko, 128, kSboP), make_desc(vh + ko, 128, kSboP), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          umma_bf16(tmem + kTmemO, make_desc(p_hi + ko, 128, kSboP), make_desc(vl + ko, 128, kSboP), idesc_o, 1);
          umma_bf16(tmem + kTmemO, make_desc(p_lo + ko, 128, kSboP), make_desc(vh + ko, 128, kSboP), idesc_o, 1);
        }
        umma_commit(&sm.kv_empty[st]);
      }
      umma_commit(&sm.o_full);
    }
  } else {
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch
    const int r = quarter * 32 + lane;            // row inside the tile
    const int row = qt * kTileQ + r;
    const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);
    const unsigned char* kmask = prm.key_mask ? prm.key_mask + (size_t)(bh % prm.n_mask_rows) * prm.Lk : nullptr;
    const unsigned char* fmask = (prm.full_mask && row < prm.Lq) ? prm.full_mask + ((size_t)(bh % prm.n_mask_rows) * prm.Lq + row) * prm.Lk : nullptr;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nKt; ++j) {
      mbar_wait(&sm.s_full, j & 1);
      tc_fence_after();
      const int key0 = j * kTileK;
      // pass A: row maximum of the (masked) scores
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float s[32];
        tmem_ld32(t_row + c * 32, s);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int key = key0 + c * 32 + i;
          bool masked = key >= prm.Lk;
          if (!masked && kmask) masked = kmask[key] != 0;
          if (!masked && fmask) masked = fmask[key] != 0;
          if (!masked) m_tile = fmaxf(m_tile, s[i]);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      // pass B: p = exp2(s - m), row sum, bf16 hi/lo -> smem (A operand of the PV MMA)
      float l_tile = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float s[32];
        tmem_ld32(t_row + c * 32, s);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi4[4], lo4[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float pv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = g * 8 + e + u;
              const int key = key0 + c * 32 + i;
              bool masked = key >= prm.Lk;
              if (!masked && kmask) masked = kmask[key] != 0;
              if (!masked && fmask) masked = fmask[key] != 0;
              pv[u] = masked ? 0.f : exp2f(s[i] - m_safe);
              l_tile += pv[u];
            }
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(pv[0], h0, l0); split_bf16(pv[1], h1, l1);
            hi4[e >> 1] = pack2(h0, h1); lo4[e >> 1] = pack2(l0, l1);
          }
          const int off = kmajor_off(r, c * 32 + g * 8, kSboP);
          *reinterpret_cast<uint4*>(sm.p_hi + off) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
          *reinterpret_cast<uint4*>(sm.p_lo + off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
        }
      }
      l_run = l_run * alpha + l_tile;
      // rescale the running output (previous PV MMAs are complete: s_full was committed after them)
      // tcgen05.ld/st are warp-collective (.sync.aligned): the branch must be warp-uniform
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        float o[16];
        tmem_ld16(t_row + kTmemO, o);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= alpha;
        tmem_st16(t_row + kTmemO, o);
      }
      m_run = m_new;
      fence_proxy_async();   // generic-proxy smem writes (P) -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(&sm.p_full);
    }
    // epilogue
    mbar_wait(&sm.o_full, 0);
    tc_fence_after();
    float o[16];
    tmem_ld16(t_row + kTmemO, o);
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
                                  int Lq, int Lk, int d, float qscale, float* __restrict__ attn) {
  const int bh = blockIdx.z, i = blockIdx.y;
  const float* qi = q + ((size_t)bh * Lq + i) * d;
  const float l2 = lse2[(size_t)bh * Lq + i];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Lk; j += gridDim.x * blockDim.x) {
    const float* kj = k + ((size_t)bh * Lk + j) * d;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s = fmaf(qi[c] * qscale, kj[c], s);
    bool masked = false;
    if (key_mask) masked = key_mask[(size_t)(bh % n_mask_rows) * Lk + j] != 0;
    if (!masked && full_mask) masked = full_mask[((size_t)(bh % n_mask_rows) * Lq + i) * Lk + j] != 0;
    attn[((size_t)bh * Lq + i) * Lk + j] = masked ? 0.f : exp2f(s - l2);
  }
}

// ---------------------------------------------------------------------------- backward (CUDA cores, fp32, recompute)
// One CTA = 64 keys of one head; loops over all queries in chunks of 64 staged in shared memory.
//   p_ij = exp2(s_ij - lse2_i);  dp_ij = dO_i . v_j;  ds_ij = p_ij (dp_ij - delta_i)
//   dV_j += p_ij dO_i;  dK_j += ds_ij q_i / T;  dQ_i += ds_ij k_j / T  (atomic, fp32)
constexpr int kBwdK = 64, kBwdQ = 64, kBwdThreads = 256;

struct BwdParams {
  const float* q; const float* k; const float* v; const float* dout; const float* out; const float* lse2;
  const unsigned char* key_mask; const unsigned char* full_mask; int n_mask_rows;
  float* dq; float* dk; float* dv;
  int BH, Lq, Lk, d, dv_dim;
  float qscale;   // log2(e)/T
  float inv_t;    // 1/T
};

__global__ void attn_delta_kernel(const float* __restrict__ dout, const float* __restrict__ out, int rows, int dv, float* __restrict__ delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float s = 0.f;
  for (int c = 0; c < dv; ++c) s = fmaf(dout[(size_t)i * dv + c], out[(size_t)i * dv + c], s);
  delta[i] = s;
}

__global__ void __launch_bounds__(kBwdThreads) attn_bwd_kernel(const BwdParams prm, const float* __restrict__ delta) {
  __shared__ float s_k[kBwdK][kDP + 1], s_v[kBwdK][kDP + 1];
  __shared__ float s_q[kBwdQ][kDP + 1], s_do[kBwdQ][kDP + 1];
  __shared__ float s_l[kBwdQ], s_dl[kBwdQ];
  __shared__ float s_dq[kBwdQ][kDP + 1];
  const int bh = blockIdx.y, k0 = blockIdx.x * kBwdK;
  const int d = prm.d, dvd = prm.dv_dim;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < kBwdK * kDP; idx += kBwdThreads) {
    const int r = idx / kDP, c = idx % kDP;
    const int key = k0 + r;
    s_k[r][c] = (key < prm.Lk && c < d) ? prm.k[((size_t)bh * prm.Lk + key) * d + c] : 0.f;
    s_v[r][c] = (key < prm.Lk && c < dvd) ? prm.v[((size_t)bh * prm.Lk + key) * dvd + c] : 0.f;
  }
  // thread -> key jj = tid % 64, query sub-lane qs = tid / 64 (4 query groups interleaved)
  const int jj = tid % kBwdK, qs = tid / kBwdK;
  const int key = k0 + jj;
  float dk_acc[kDP], dv_acc[kDP];
#pragma unroll
  for (int c = 0; c < kDP; ++c) { dk_acc[c] = 0.f; dv_acc[c] = 0.f; }
  const unsigned char* kmask = prm.key_mask ? prm.key_mask + (size_t)(bh % prm.n_mask_rows) * prm.Lk : nullptr;
  const bool key_ok = key < prm.Lk && !(kmask && kmask[key] != 0);
  __syncthreads();
  float kreg[kDP], vreg[kDP];
#pragma unroll
  for (int c = 0; c < kDP; ++c) { kreg[c] = s_k[jj][c]; vreg[c] = s_v[jj][c]; }

  for (int q0 = 0; q0 < prm.Lq; q0 += kBwdQ) {
    __syncthreads();
    for (int idx = tid; idx < kBwdQ * kDP; idx += kBwdThreads) {
      const int r = idx / kDP, c = idx % kDP;
      const int qi = q0 + r;
      s_q[r][c] = (qi < prm.Lq && c < d) ? prm.q[((size_t)bh * prm.Lq + qi) * d + c] : 0.f;
      s_do[r][c] = (qi < prm.Lq && c < dvd) ? prm.dout[((size_t)bh * prm.Lq + qi) * dvd + c] : 0.f;
      s_dq[r][c] = 0.f;
    }
    for (int r = tid; r < kBwdQ; r += kBwdThreads) {
      const int qi = q0 + r;
      s_l[r] = (qi < prm.Lq) ? prm.lse2[(size_t)bh * prm.Lq + qi] : 0.f;
      s_dl[r] = (qi < prm.Lq) ? delta[(size_t)bh * prm.Lq + qi] : 0.f;
    }
    __syncthreads();
    for (int r = qs; r < kBwdQ; r += kBwdThreads / kBwdK) {
      const int qi = q0 + r;
      float ds = 0.f;
      if (qi < prm.Lq && key_ok) {
        bool masked = false;
        if (prm.full_mask) masked = prm.full_mask[((size_t)(bh % prm.n_mask_rows) * prm.Lq + qi) * prm.Lk + key] != 0;
        if (!masked) {
          float s = 0.f, dp = 0.f;
#pragma unroll
          for (int c = 0; c < kDP; ++c) { s = fmaf(s_q[r][c], kreg[c], s); dp = fmaf(s_do[r][c], vreg[c], dp); }
          const float p = exp2f(s * prm.qscale - s_l[r]);
          ds = p * (dp - s_dl[r]) * prm.inv_t;
#pragma unroll
          for (int c = 0; c < kDP; ++c) { dv_acc[c] = fmaf(p, s_do[r][c], dv_acc[c]); dk_acc[c] = fmaf(ds, s_q[r][c], dk_acc[c]); }
        }
      }
      // dQ_i += sum_j ds_ij k_j : reduce over the 64 keys of this CTA (two warps per query row)
#pragma unroll
      for (int c = 0; c < kDP; ++c) {
        float t = ds * kreg[c];
        t = warp_sum(t);
        if ((tid & 31) == 0 && c < d) atomicAdd(&s_dq[r][c], t);
      }
    }
    __syncthreads();
    for (int idx = tid; idx < kBwdQ * d; idx += kBwdThreads) {
      const int r = idx / d, c = idx % d;
      const int qi = q0 + r;
      if (qi < prm.Lq) atomicAdd(prm.dq + ((size_t)bh * prm.Lq + qi) * d + c, s_dq[r][c]);
    }
  }
  // combine the 4 query groups of each key
  __syncthreads();
  float(*s_red)[kDP + 1] = s_q;  // reuse [64][17]
  for (int pass = 0; pass < 2; ++pass) {
    float* acc = pass == 0 ? dk_acc : dv_acc;
    for (int g = 1; g < kBwdThreads / kBwdK; ++g) {
      __syncthreads();
      if (qs == g)
        for (int c = 0; c < kDP; ++c) s_red[jj][c] = acc[c];
      __syncthreads();
      if (qs == 0)
        for (int c = 0; c < kDP; ++c) acc[c] += s_red[jj][c];
    }
    if (qs == 0 && key < prm.Lk) {
      if (pass == 0) for (int c = 0; c < d; ++c) prm.dk[((size_t)bh * prm.Lk + key) * d + c] = acc[c];
      else for (int c = 0; c < dvd; ++c) prm.dv[((size_t)bh * prm.Lk + key) * dvd + c] = acc[c];
    }
  }
}

int check_attn(int BH, int Lq, int Lk, int d, int dv) {
  ISA_CHECK_ARG(BH > 0 && Lq > 0 && Lk > 0, "attention: non-positive size (BH=%d Lq=%d Lk=%d)", BH, Lq, Lk);
  ISA_CHECK_ARG(d > 0 && d <= kDP && dv > 0 && dv <= kDP, "attention: head widths must be <= %d (d_k=%d, d_v=%d)", kDP, d, dv);
  ISA_CHECK_ARG(BH <= 65535, "attention: n_head*batch=%d > 65535", BH);
  return ISA_OK;
}

}  // namespace

extern "C" {

size_t isa_attention_workspace_bytes(int BH, int Lq, int Lk) {
  if (BH <= 0 || Lq <= 0 || Lk <= 0) return 0;
  const size_t nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  return isa_align_up((size_t)BH * nQt * kQTileBytes, 1024) + isa_align_up((size_t)BH * nKt * kKvTileBytes, 1024) +
         isa_align_up(sizeof(float) * (size_t)BH * Lq, 1024);
}

// q [BH][Lq][d], k [BH][Lk][d], v [BH][Lk][dv] fp32; temperature T (= sqrt(d_k) in the reference);
// key_mask [n_mask_rows][Lk] u8 (1 = masked, broadcast over queries; row = bh % n_mask_rows) and/or
// full_mask [n_mask_rows][Lq][Lk] u8; out [BH][Lq][dv]; lse2 [BH][Lq] (may be NULL).
int isa_attention_fwd(const float* q, const float* k, const float* v, int BH, int Lq, int Lk, int d, int dv, float temperature,
                      const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows,
                      float* out, float* lse2, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, dv);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && v && out && workspace, "attention_fwd: null pointer");
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
  fp.qblk = qblk; fp.kvblk = kvblk; fp.key_mask = key_mask; fp.full_mask = full_mask; fp.n_mask_rows = n_mask_rows > 0 ? n_mask_rows : 1;
  fp.out = out; fp.lse2 = lse2; fp.BH = BH; fp.Lq = Lq; fp.Lk = Lk; fp.dv = dv; fp.nQt = nQt; fp.nKt = nKt;
  const size_t smem = sizeof(FwdSmem) + 1024;
  ISA_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<<<dim3(nQt, BH), kFwdThreads, smem, stream>>>(fp);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// attn [BH][Lq][Lk] = softmax probabilities (0 where masked) from the saved lse2.
int isa_attention_probs(const float* q, const float* k, const float* lse2, int BH, int Lq, int Lk, int d, float temperature,
                        const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows, float* attn, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, 1);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && lse2 && attn, "attention_probs: null pointer");
  ISA_CHECK_ARG(Lq <= 65535, "attention_probs: Lq=%d > 65535", Lq);
  int gx = (Lk + 255) / 256;
  if (gx > 64) gx = 64;
  attn_probs_kernel<<<dim3(gx, Lq, BH), 256, 0, stream>>>(q, k, lse2, key_mask, full_mask, n_mask_rows > 0 ? n_mask_rows : 1, Lq, Lk, d,
                                                         1.4426950408889634f / temperature, attn);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// dq/dk/dv from (q,k,v,out,dout,lse2); dq is zeroed here.  workspace: same size as forward (its tail holds delta).
int isa_attention_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout, const float* lse2,
                      int BH, int Lq, int Lk, int d, int dv, float temperature,
                      const unsigned char* key_mask, const unsigned char* full_mask, int n_mask_rows,
                      float* dq, float* dk, float* dvv, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_attn(BH, Lq, Lk, d, dv);
  if (rc) return rc;
  ISA_CHECK_ARG(q && k && v && out && dout && lse2 && dq && dk && dvv && workspace, "attention_bwd: null pointer");
  if (workspace_bytes < isa_attention_workspace_bytes(BH, Lq, Lk)) {
    isa_set_error("attention_bwd: workspace too small");
    return ISA_ERR_WORKSPACE;
  }
  const size_t nQt = (Lq + kTileQ - 1) / kTileQ, nKt = (Lk + kTileK - 1) / kTileK;
  float* delta = (float*)((unsigned char*)workspace + isa_align_up((size_t)BH * nQt * kQTileBytes, 1024) + isa_align_up((size_t)BH * nKt * kKvTileBytes, 1024));
  const int rows = BH * Lq;
  attn_delta_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(dout, out, rows, dv, delta);
  ISA_CUDA(cudaMemsetAsync(dq, 0, sizeof(float) * (size_t)BH * Lq * d, stream));
  BwdParams bp;
  bp.q = q; bp.k = k; bp.v = v; bp.dout = dout; bp.out = out; bp.lse2 = lse2;
  bp.key_mask = key_mask; bp.full_mask = full_mask; bp.n_mask_rows = n_mask_rows > 0 ? n_mask_rows : 1;
  bp.dq = dq; bp.dk = dk; bp.dv = dvv; bp.BH = BH; bp.Lq = Lq; bp.Lk = Lk; bp.d = d; bp.dv_dim = dv;
  bp.qscale = 1.4426950408889634f / temperature; bp.inv_t = 1.f / temperature;
  attn_bwd_kernel<<<dim3((Lk + kBwdK - 1) / kBwdK, BH), kBwdThreads, 0, stream>>>(bp, delta);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
