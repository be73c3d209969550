// ReNet projection GEMMs on tcgen05 with the fp32 -> bf16 hi/lo operand split fused into the producer warps.
//
// The GEMMs around the GRU scan (reference contract /root/reference/code/lib/archs/modules/README.md:225-256; the
// arithmetic is nn.GRU's x W_ih^T and autograd's dG W_ih, dG^T [x | h_prev | 1]) ran as isa_split_bf16x3 (4 B read +
// 6 B written per element) followed by K-concatenated cuBLAS GEMMs over the split copies: ~16 B of HBM traffic per
// operand element against 4 B algorithmic.  Here ONE kernel reads the fp32 operands once:
//   * 16 producer warps load 8-element runs (32 B, a full sector) of both operands straight from global memory, split
//     them in registers into bf16 hi + lo parts (a = hi + lo to ~16 mantissa bits) and store 16-byte rows of the
//     canonical no-swizzle UMMA core matrices into a 3-stage shared-memory ring (conflict free: a quarter warp writes
//     128 contiguous bytes).  The operand may be K-contiguous (rows = M/N index) or MN-contiguous (rows = K index): the
//     store picks the K-major or MN-major core-matrix layout and the instruction descriptor's a_major / b_major bits
//     do the transposition -- no operand is ever transposed or copied in global memory;
//   * warp 0 issues (one elected lane), per 16-deep k step, a_hi b_hi + a_hi b_lo + a_lo b_hi (M = 128, N <= 256, fp32 accumulation
//     in tensor memory) -- fp32-level accuracy (error ~2^-16 relative) like the rest of the hot path; a CTA owns 256 rows
//     (two accumulators, all 512 TMEM columns) so every B stage is converted / fetched once for both halves;
//   * mbarriers: full[stage] <- producer arrivals (+ the bulk copy's bytes), empty[stage] <- tcgen05.commit; 3 stages.
// MN-major operands are concatenations of up to four column segments, each optionally read with a row shift and
// zero-filled where the shifted row leaves its sweep (h_{t-1} of a direction straight from the forward output) or
// all ones (the bias-gradient column), so [dgx | dghn]^T [x | h_prev(0) | h_prev(1) | 1] needs no materialised operand.
// The weight gradient (K = all tokens) is split over blockIdx.z; partials are folded in a fixed order (deterministic).
#include "isa_common.cuh"
#include "isa_tcgen05.cuh"

namespace {

constexpr int kBM = 256;                                 // two 128-row accumulators per CTA share every B stage
constexpr int kBNMax = 256;
constexpr int kBK = 32;
constexpr int kStages = 3;
constexpr int kProducerThreads = 512;
constexpr int kThreadsTotal = kProducerThreads;          // 16 warps: 4 per SM sub-partition -> 128 registers per thread (warp 0 also issues the MMAs)
constexpr int kTileA = kBM * kBK * 2;                    // bytes of one bf16 part of the A stage (16 KB)
constexpr int kTileB = kBNMax * kBK * 2;                 // 16 KB
constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;     // hi + lo of both operands: 64 KB
constexpr int kGroupStride = (kBK / 8) * 128;            // 512 B between adjacent 8-wide M/N groups of a stage tile
constexpr int kMaxSeg = 4;

struct Seg {
  const float* ptr;      // element (row r, col c) at ptr[r * ld + c]
  long long ld;
  long long shift;       // MN-major only: read row r + shift ...
  int cols;              // real columns (K-major: the K extent; MN-major: M/N columns of this segment)
  int v0;                // MN-major: first virtual column of the segment (multiple of 8)
  int pos_div, pos_mod, pos_step;   // ... zero where ((r / pos_div) % pos_mod) + pos_step leaves [0, pos_mod)
  int ones;              // the segment reads as all ones
};

struct Operand {
  Seg seg[kMaxSeg];
  int nseg;
  int mn_major;          // 0: rows = M/N index, K contiguous (single segment); 1: rows = K index, M/N contiguous
  int extent;            // M/N extent (K-major: rows; MN-major: virtual columns, multiple of 8)
};

struct GemmParams {
  Operand a, b;
  long long K;           // reduction length
  long long k_per_split; // multiple of kBK
  float* out;            // [splits][M_out_rows][ldc] (splits == 1: the result itself)
  long long ldc;
  long long out_rows;    // rows of one split's output (stores are bounded by it)
  int out_cols;          // stores are bounded by it
  int bn;                // N tile (multiple of 16, <= 256)
  const unsigned char* b_packed;   // packed B operand (pack_operand_kernel) or nullptr
  int n_iter_total;      // k blocks of the whole K (packed addressing)
};

// ---- producer side: every thread owns a fixed set of 16-byte core-matrix rows ("chunks": 8 consecutive elements along
// the contiguous global dimension) of each operand stage; its plan (global pointer, shared-memory offset, validity) is
// computed once and advanced by one k block per iteration, so the loop body is loads, F2FP splits and two 16-byte stores.
struct Chunk {
  const float* p;     // first element of the chunk at iteration 0 (nullptr: reads as zero for the whole tile)
  long long step;     // elements to advance per iteration (K-major: kBK; MN-major: kBK * ld)
  int off;            // byte offset inside a bf16 stage tile
  int k0;             // K-major: first k of the chunk at iteration 0; MN-major: the chunk's k row at iteration 0
  int tail;           // MN-major: valid columns (1..8); K-major: unused
  int seg;            // MN-major: segment index (position test / ones), -1 = plain
};

__device__ __forceinline__ void plan_chunk(const Operand& op, long long mn_base, int groups, long long k_lo, int q, Chunk& c) {
  c.p = nullptr; c.step = 0; c.off = 0; c.k0 = 0; c.tail = 8; c.seg = -1;
  if (op.mn_major) {
    // a warp = 8 k rows x 4 adjacent M/N groups (128 contiguous bytes per global row)
    const int k8 = q & 7, mgl = (q >> 3) & 3, rest = q >> 5, kg = rest & 3, mg = (rest >> 2) * 4 + mgl;
    if (mg >= groups) { c.off = -1; return; }
    c.off = mg * kGroupStride + kg * 128 + k8 * 16;
    c.k0 = (int)k_lo + kg * 8 + k8;
    const int mn0 = (int)mn_base + mg * 8;
    if (mn0 >= op.extent) return;
    int si = 0;
#pragma unroll
    for (int j = 1; j < kMaxSeg; ++j)
      if (j < op.nseg && mn0 >= op.seg[j].v0) si = j;
    const Seg& sg = op.seg[si];
    const int col = mn0 - sg.v0;
    if (col >= sg.cols) return;
    c.tail = sg.cols - col < 8 ? sg.cols - col : 8;
    c.seg = (sg.ones || sg.pos_step != 0) ? si : -1;
    c.step = (long long)kBK * sg.ld;
    c.p = sg.ones ? reinterpret_cast<const float*>(&op) /* non-null marker */ : sg.ptr + ((long long)c.k0 + sg.shift) * sg.ld + col;
  } else {
    // a warp = one 8-row group, its 4 k-chunks (128 contiguous bytes per global row)
    const int r8 = q & 7, kc = (q >> 3) & 3, rg = q >> 5;
    if (rg >= groups) { c.off = -1; return; }
    c.off = rg * kGroupStride + kc * 128 + r8 * 16;
    c.k0 = (int)k_lo + kc * 8;
    const long long mn = mn_base + rg * 8 + r8;
    if (mn >= op.extent) return;
    c.step = kBK;
    c.p = op.seg[0].ptr + mn * op.seg[0].ld + c.k0;
  }
}

__device__ __forceinline__ void load_chunk(const Operand& op, const Chunk& c, int it, int k_hi, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (c.p == nullptr) return;
  const int k = c.k0 + it * kBK;
  if (k >= k_hi) return;
  const float* p = c.p + (long long)it * c.step;
  if (op.mn_major) {
    if (c.seg >= 0) {
      const Seg& sg = op.seg[c.seg];
      if (sg.ones) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = i < c.tail ? 1.f : 0.f;
        return;
      }
      const int pos = (int)(((unsigned)k / (unsigned)sg.pos_div) % (unsigned)sg.pos_mod) + sg.pos_step;
      if (pos < 0 || pos >= sg.pos_mod) return;
    }
    if (c.tail == 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < c.tail) v[i] = __ldg(p + i);
    }
  } else {
    if (k + 8 <= k_hi) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k + i < k_hi) v[i] = __ldg(p + i);
    }
  }
}

__device__ __forceinline__ void store_split8(unsigned char* hi_tile, unsigned char* lo_tile, int off, const float (&v)[8]) {
  uint4 h, l;
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(hi_tile + off) = h;
  *reinterpret_cast<uint4*>(lo_tile + off) = l;
}

// Packs a (weight) operand once into the exact per-(N tile, k block) shared-memory image [hi tile | lo tile] so the GEMM
// brings a stage of it with two bulk copies instead of re-splitting it for every M tile.  grid = (k blocks, N tiles).
__global__ void __launch_bounds__(kProducerThreads) pack_operand_kernel(const __grid_constant__ Operand op, int bn, int k_total, unsigned char* __restrict__ dst) {
  const int it = blockIdx.x, nt = blockIdx.y;
  const int n0 = nt * bn;
  int width = op.extent - n0;
  width = width > bn ? bn : ((width + 15) & ~15);
  unsigned char* hi = dst + ((size_t)nt * gridDim.x + it) * 2 * (size_t)bn * (kBK * 2);
  unsigned char* lo = hi + (size_t)bn * (kBK * 2);
#pragma unroll 1
  for (int i = 0; i < kBNMax * (kBK / 8) / kProducerThreads; ++i) {
    Chunk c;
    plan_chunk(op, n0, width / 8, 0, threadIdx.x + i * kProducerThreads, c);
    if (c.off < 0) continue;
    float v[8];
    load_chunk(op, c, it, k_total, v);
    store_split8(hi, lo, c.off, v);
  }
}

template <bool PACKED_B>
__global__ void __launch_bounds__(kThreadsTotal, 1) proj_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw_g[];
  unsigned char* sm = smem_raw_g + ((128u - (smem_u32(smem_raw_g) & 127u)) & 127u);
  __shared__ uint64_t s_full[kStages], s_empty[kStages], s_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long m0 = (long long)blockIdx.y * kBM;
  const int n0 = blockIdx.x * p.bn;
  const long long k_lo = (long long)blockIdx.z * p.k_per_split;
  const int k_hi = (int)((k_lo + p.k_per_split < p.K) ? k_lo + p.k_per_split : p.K);
  const int n_iter = (int)((k_hi - k_lo + kBK - 1) / kBK);
  int bn = p.b.extent - n0;
  bn = bn > p.bn ? p.bn : ((bn + 15) & ~15);
  const int n_acc = (p.a.extent - m0 > 128) ? 2 : 1;       // the second 128-row half may be empty (last M tile)

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], kProducerThreads / 32 + (PACKED_B ? 1 : 0)); mbar_init(&s_empty[s], 1); }
    mbar_init(&s_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  {
    // ------------------------------------------------------------ producers (warp 0 also issues the MMAs of every stage it has helped fill)
    const uint32_t el = elect_one();
    const uint32_t idesc = make_idesc(128, bn) | ((uint32_t)(p.a.mn_major ? 1 : 0) << 15) | ((uint32_t)(p.b.mn_major ? 1 : 0) << 16);
    constexpr int NA = kBM * (kBK / 8) / kProducerThreads;                       // 2 chunks of A per thread
    constexpr int NB = PACKED_B ? 1 : kBNMax * (kBK / 8) / kProducerThreads;     // 2 chunks of B per thread
    Chunk ca[NA], cb[NB];
    float va[NA][8], vb[NB][8], na[NA][8], nb_[NB][8];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      plan_chunk(p.a, m0, n_acc * 16, k_lo, tid + i * kProducerThreads, ca[i]);
      if (ca[i].off >= 0) load_chunk(p.a, ca[i], 0, k_hi, va[i]);
    }
    if (!PACKED_B) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        plan_chunk(p.b, n0, bn / 8, k_lo, tid + i * kProducerThreads, cb[i]);
        if (cb[i].off >= 0) load_chunk(p.b, cb[i], 0, k_hi, vb[i]);
      }
    }
    const size_t pk_tile = (size_t)p.bn * (kBK * 2);
    const int lane_p = tid & 31;
    // one k block: `cur` holds its operand values (loaded one block earlier), `nxt` receives the next block's.  The
    // loads are issued first and stay in flight for the whole block (barrier wait, split, stores); the caller
    // alternates the two register sets, so nothing waits on them before the next block's split
    auto k_block = [&](int it, float (&cur_a)[NA][8], float (&cur_b)[NB][8], float (&nxt_a)[NA][8], float (&nxt_b)[NB][8]) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      if (it + 1 < n_iter) {
#pragma unroll
        for (int i = 0; i < NA; ++i)
          if (ca[i].off >= 0) load_chunk(p.a, ca[i], it + 1, k_hi, nxt_a[i]);
        if (!PACKED_B) {
#pragma unroll
          for (int i = 0; i < NB; ++i)
            if (cb[i].off >= 0) load_chunk(p.b, cb[i], it + 1, k_hi, nxt_b[i]);
        }
      }
      if (lane_p == 0) mbar_wait(&s_empty[s], ph ^ 1u);
      __syncwarp();
      unsigned char* st = sm + s * kStageBytes;
      if (PACKED_B && tid == 0) {
        const unsigned char* src = p.b_packed + ((size_t)blockIdx.x * p.n_iter_total + (size_t)(k_lo / kBK) + it) * 2 * pk_tile;
        const uint32_t bytes = (uint32_t)bn * (kBK * 2);
        mbar_arrive_expect_tx(&s_full[s], 2 * bytes);
        tma_bulk_g2s(st + 2 * kTileA, src, bytes, &s_full[s]);
        tma_bulk_g2s(st + 2 * kTileA + kTileB, src + pk_tile, bytes, &s_full[s]);
      }
#pragma unroll
      for (int i = 0; i < NA; ++i)
        if (ca[i].off >= 0) store_split8(st, st + kTileA, ca[i].off, cur_a[i]);
      if (!PACKED_B) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
          if (cb[i].off >= 0) store_split8(st + 2 * kTileA, st + 2 * kTileA + kTileB, cb[i].off, cur_b[i]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane_p == 0) mbar_arrive(&s_full[s]);          // one arrival per producer warp
      if (warp == 0) {
        // warp-uniform MMA issue, one elected lane: a_hi b_hi + a_hi b_lo + a_lo b_hi per 16-deep k step and accumulator
        mbar_wait(&s_full[s], ph);
        tc_fence_after();
        const uint32_t base = smem_u32(st);
        const uint64_t b_hi = make_desc(base + 2 * kTileA, 128, kGroupStride), b_lo = make_desc(base + 2 * kTileA + kTileB, 128, kGroupStride);
        for (int acc = 0; acc < n_acc; ++acc) {
          const uint32_t a_off = (uint32_t)acc * (16 * kGroupStride);           // 128 rows = 16 groups
          const uint64_t a_hi = make_desc(base + a_off, 128, kGroupStride), a_lo = make_desc(base + kTileA + a_off, 128, kGroupStride);
          const uint32_t d = tmem + (uint32_t)acc * 256u;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t ko = (uint64_t)((kk * 256) >> 4);
            umma_bf16_e(el, d, a_hi + ko, b_hi + ko, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            umma_bf16_e(el, d, a_hi + ko, b_lo + ko, idesc, 1u);
            umma_bf16_e(el, d, a_lo + ko, b_hi + ko, idesc, 1u);
          }
        }
        umma_commit_e(el, &s_empty[s]);
        if (it + 1 == n_iter) umma_commit_e(el, &s_done);
      }
    };
    for (int it = 0; it < n_iter; it += 2) {
      k_block(it, va, vb, na, nb_);
      if (it + 1 < n_iter) k_block(it + 1, na, nb_, va, vb);
    }
    // ------------------------------------------------------------ epilogue: warp w reads accumulator w / 4, TMEM lane quarter w % 4
    if (warp < 4 * n_acc) {
      mbar_wait(&s_done, 0);
      tc_fence_after();
      const int lane = tid & 31;
      const int acc = warp >> 2, quarter = warp & 3;
      const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * 256u;
      // all MMAs have completed: the stage ring is free, each warp transposes 32 x 32 blocks through 4.5 KB of it so that a
      // quarter warp writes 128 contiguous bytes of one output row
      float* stg = reinterpret_cast<float*>(sm) + warp * (32 * 36);
      const long long row_base = m0 + acc * 128 + quarter * 32;
      float* obase = p.out + ((long long)blockIdx.z * p.out_rows + row_base) * p.ldc + n0;
      for (int c0 = 0; c0 < bn; c0 += 32) {
        uint32_t d[32];
        if (bn - c0 >= 32) {
          tmem_ld32_issue(t_row + c0, d);
          tmem_ld32_wait(d);
        } else {   // bn is a multiple of 16
          uint32_t e[16];
          tmem_ld16_issue(t_row + c0, e);
          tmem_ld16_wait(e);
#pragma unroll
          for (int i = 0; i < 16; ++i) { d[i] = e[i]; d[i + 16] = 0u; }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(stg + lane * 36 + i) = make_uint4(d[i], d[i + 1], d[i + 2], d[i + 3]);
        __syncwarp();
        int lim = p.out_cols - (n0 + c0);
        if (lim > bn - c0) lim = bn - c0;      // the tile's own columns only
        const int col = (lane & 7) * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = j * 4 + (lane >> 3);
          const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 36 + col);
          if (row_base + r < p.out_rows) {
            float* o = obase + (long long)r * p.ldc + c0 + col;
            if (col + 4 <= lim) *reinterpret_cast<uint4*>(o) = v;
            else {
              if (col + 0 < lim) o[0] = __uint_as_float(v.x);
              if (col + 1 < lim) o[1] = __uint_as_float(v.y);
              if (col + 2 < lim) o[2] = __uint_as_float(v.z);
              if (col + 3 < lim) o[3] = __uint_as_float(v.w);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Folds the split-K partials of the weight-gradient product in a fixed order and scatters the blocks:
//   rows    [0, 6n)  = dgx columns (direction-major: d * 3n + gate * n + unit),  [6n, 8n) = dghn (d * n + unit)
//   columns [0, cin) = x,  [cb0, cb0 + n) = h_prev dir 0,  [cb1, cb1 + n) = h_prev dir 1,  cone = ones
__global__ void __launch_bounds__(256) renet_wgrad_reduce_kernel(const float* __restrict__ part, int splits, long long rows_pad, long long ldc,
                                                                 int n, int cin, int cb0, int cb1, int cone, float* __restrict__ dw_ih,
                                                                 float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh) {
  const int n_ih = 6 * n * cin, n_hh = 2 * 3 * n * n, n_b = 6 * n;
  const int total = n_ih + n_hh + 2 * n_b;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r, c;
    float* dst;
    if (i < n_ih) { r = i / cin; c = i % cin; dst = dw_ih + i; }
    else if (i < n_ih + n_hh) {
      const int j = i - n_ih;                 // dw_hh[d][g][u]: d = j / (3 n n), g = row within 3n, u = column
      const int d = j / (3 * n * n), g = (j / n) % (3 * n), u = j % n;
      r = g < 2 * n ? d * 3 * n + g : 6 * n + d * n + (g - 2 * n);
      c = (d == 0 ? cb0 : cb1) + u;
      dst = dw_hh + j;
    } else if (i < n_ih + n_hh + n_b) { r = i - n_ih - n_hh; c = cone; dst = db_ih + r; }
    else {
      const int j = i - n_ih - n_hh - n_b;    // db_hh[d][g]
      const int d = j / (3 * n), g = j % (3 * n);
      r = g < 2 * n ? d * 3 * n + g : 6 * n + d * n + (g - 2 * n);
      c = cone;
      dst = db_hh + j;
    }
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[((long long)s * rows_pad + r) * ldc + c];
    *dst = acc;
  }
}

Seg plain_seg(const float* ptr, long long ld, int cols, int v0) {
  Seg s;
  s.ptr = ptr; s.ld = ld; s.shift = 0; s.cols = cols; s.v0 = v0; s.pos_div = 1; s.pos_mod = 1; s.pos_step = 0; s.ones = 0;
  return s;
}

int pick_bn(int extent) {
  // fewest N tiles, then the least padding (multiples of 16)
  const int tiles = (extent + kBNMax - 1) / kBNMax;
  int bn = (extent + tiles - 1) / tiles;
  return (bn + 15) & ~15;
}

int pick_bn_fwd(int extent) { return pick_bn(extent); }

int launch_gemm(const GemmParams& p, long long M, int splits, cudaStream_t stream) {
  static thread_local bool attr_set[16] = {false};
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  const size_t smem = (size_t)kStages * kStageBytes + 128;
  if (di.device >= 0 && di.device < 16 && !attr_set[di.device]) {
    ISA_CUDA(cudaFuncSetAttribute(proj_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ISA_CUDA(cudaFuncSetAttribute(proj_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[di.device] = true;
  }
  dim3 grid((unsigned)((p.b.extent + p.bn - 1) / p.bn), (unsigned)((M + kBM - 1) / kBM), (unsigned)splits);
  if (p.b_packed) proj_gemm_kernel<true><<<grid, kThreadsTotal, smem, stream>>>(p);
  else proj_gemm_kernel<false><<<grid, kThreadsTotal, smem, stream>>>(p);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// packs operand B of `p` into `workspace` and points the GEMM at it
int pack_b(GemmParams& p, void* workspace, size_t workspace_bytes, const char* who, cudaStream_t stream) {
  const int n_tiles = (p.b.extent + p.bn - 1) / p.bn;
  const int k_blocks = (int)((p.K + kBK - 1) / kBK);
  const size_t need = (size_t)n_tiles * k_blocks * 2 * p.bn * (kBK * 2);
  if (workspace == nullptr || workspace_bytes < need) {
    isa_set_error("%s: workspace %zu < %zu", who, workspace_bytes, need);
    return ISA_ERR_WORKSPACE;
  }
  pack_operand_kernel<<<dim3(k_blocks, n_tiles), kProducerThreads, 0, stream>>>(p.b, p.bn, (int)p.K, reinterpret_cast<unsigned char*>(workspace));
  ISA_CUDA(cudaGetLastError());
  p.b_packed = reinterpret_cast<const unsigned char*>(workspace);
  p.n_iter_total = k_blocks;
  return ISA_OK;
}

size_t pack_bytes(int extent, int K) {
  const int bn = pick_bn_fwd(extent);
  const int n_tiles = (extent + bn - 1) / bn;
  const int k_blocks = (K + kBK - 1) / kBK;
  return (size_t)n_tiles * k_blocks * 2 * bn * (kBK * 2);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

size_t isa_renet_proj_workspace_bytes(int cin, int n_out) {
  if (cin <= 0 || n_out <= 0) return 0;
  const size_t a = pack_bytes(n_out, cin);               // fwd: B = w as [n_out] x K = cin
  const size_t b = pack_bytes((cin + 7) & ~7, n_out);    // dx:  B = w as [cin] x K = n_out
  return a > b ? a : b;
}

// gx[tokens][n_out] = x[tokens][cin] w[n_out][cin]^T
int isa_renet_proj_fwd(const float* x, const float* w, long long tokens, int cin, int n_out, float* gx, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(x && w && gx, "renet_proj_fwd: null pointer");
  ISA_CHECK_ARG(tokens > 0 && cin > 0 && n_out > 0 && cin % 4 == 0 && n_out % 4 == 0,
                "renet_proj_fwd: tokens=%lld cin=%d n_out=%d (widths must be positive multiples of 4)", tokens, cin, n_out);
  ISA_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(gx) && aligned16(workspace), "renet_proj_fwd: pointers must be 16-byte aligned");
  ISA_CHECK_ARG(tokens < 0x7fffffff, "renet_proj_fwd: too many tokens");
  GemmParams p;
  p.a.nseg = 1; p.a.mn_major = 0; p.a.extent = (int)tokens;
  p.a.seg[0] = plain_seg(x, cin, cin, 0);
  p.b.nseg = 1; p.b.mn_major = 0; p.b.extent = n_out;
  p.b.seg[0] = plain_seg(w, cin, cin, 0);
  p.K = cin; p.k_per_split = ((cin + kBK - 1) / kBK) * kBK;
  p.out = gx; p.ldc = n_out; p.out_rows = tokens; p.out_cols = n_out;
  p.bn = pick_bn(n_out);
  int rc = pack_b(p, workspace, workspace_bytes, "renet_proj_fwd", stream);
  if (rc) return rc;
  return launch_gemm(p, tokens, 1, stream);
}

// dx[tokens][cin] = dg[tokens][n_out] w[n_out][cin]
int isa_renet_proj_dx(const float* dg, const float* w, long long tokens, int n_out, int cin, float* dx, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(dg && w && dx, "renet_proj_dx: null pointer");
  ISA_CHECK_ARG(tokens > 0 && tokens < 0x7fffffff && cin > 0 && n_out > 0 && cin % 4 == 0 && n_out % 4 == 0,
                "renet_proj_dx: tokens=%lld cin=%d n_out=%d (widths must be positive multiples of 4)", tokens, cin, n_out);
  ISA_CHECK_ARG(aligned16(dg) && aligned16(w) && aligned16(dx) && aligned16(workspace), "renet_proj_dx: pointers must be 16-byte aligned");
  GemmParams p;
  p.a.nseg = 1; p.a.mn_major = 0; p.a.extent = (int)tokens;
  p.a.seg[0] = plain_seg(dg, n_out, n_out, 0);
  p.b.nseg = 1; p.b.mn_major = 1; p.b.extent = (cin + 7) & ~7;
  p.b.seg[0] = plain_seg(w, cin, cin, 0);
  p.K = n_out; p.k_per_split = ((n_out + kBK - 1) / kBK) * kBK;
  p.out = dx; p.ldc = cin; p.out_rows = tokens; p.out_cols = cin;
  p.bn = pick_bn(p.b.extent);
  int rc = pack_b(p, workspace, workspace_bytes, "renet_proj_dx", stream);
  if (rc) return rc;
  return launch_gemm(p, tokens, 1, stream);
}

static void wgrad_geometry(int n, int cin, long long tokens, int num_sms, int* cb0, int* cb1, int* cone, int* nb, long long* rows_pad,
                           int* splits, long long* k_per) {
  const int nv = (n + 7) & ~7, cv = (cin + 7) & ~7;
  *cb0 = cv; *cb1 = cv + nv; *cone = cv + 2 * nv; *nb = cv + 2 * nv + 8;
  const long long m_tiles = (8LL * n + kBM - 1) / kBM;
  *rows_pad = m_tiles * kBM;
  const int n_tiles = (*nb + pick_bn(*nb) - 1) / pick_bn(*nb);
  long long tiles = m_tiles * n_tiles;
  int s = (int)((long long)num_sms / tiles);               // one CTA per SM, one wave
  const long long max_s = (tokens + 4 * kBK - 1) / (4 * kBK);
  if (s > max_s) s = (int)max_s;
  if (s < 1) s = 1;
  long long kp = (tokens + s - 1) / s;
  kp = (kp + kBK - 1) / kBK * kBK;
  *splits = (int)((tokens + kp - 1) / kp);
  *k_per = kp;
}

size_t isa_renet_proj_wgrad_workspace_bytes(long long tokens, int cin, int n) {
  if (tokens <= 0 || cin <= 0 || n <= 0) return 0;
  IsaDeviceInfo di;
  if (isa_device_info(&di) != ISA_OK) { di.num_sms = 148; }
  int cb0, cb1, cone, nb, splits;
  long long rows_pad, kper;
  wgrad_geometry(n, cin, tokens, di.num_sms, &cb0, &cb1, &cone, &nb, &rows_pad, &splits, &kper);
  return (size_t)splits * rows_pad * nb * sizeof(float);
}

// [dgx | dghn]^T [x | h_prev(dir 0) | h_prev(dir 1) | 1] over all tokens -> dW_ih [2][3n][cin], dW_hh [2][3n][n], db_ih [2][3n], db_hh [2][3n]
// dgx [tokens][2][3n], dghn [tokens][2][n], x [tokens][cin], out [tokens][2][n] (the forward output: h_{t-1} of direction 0 is the row
// `step` tokens back, of direction 1 `step` tokens ahead, zero outside the sweep: position = (token / pos_div) % pos_mod).
int isa_renet_proj_wgrad(const float* dgx, const float* dghn, const float* x, const float* out, long long tokens, int cin, int n,
                         long long step, int pos_div, int pos_mod, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(dgx && dghn && x && out && dw_ih && dw_hh && db_ih && db_hh && workspace, "renet_proj_wgrad: null pointer");
  ISA_CHECK_ARG(tokens > 0 && tokens < 0x7fffffff && cin > 0 && n > 0 && cin % 4 == 0 && n % 4 == 0,
                "renet_proj_wgrad: tokens=%lld cin=%d n=%d (widths must be positive multiples of 4)", tokens, cin, n);
  ISA_CHECK_ARG(pos_div > 0 && pos_mod > 0 && step > 0, "renet_proj_wgrad: step / pos_div / pos_mod must be positive");
  ISA_CHECK_ARG(aligned16(dgx) && aligned16(dghn) && aligned16(x) && aligned16(out), "renet_proj_wgrad: pointers must be 16-byte aligned");
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  int cb0, cb1, cone, nb, splits;
  long long rows_pad, kper;
  wgrad_geometry(n, cin, tokens, di.num_sms, &cb0, &cb1, &cone, &nb, &rows_pad, &splits, &kper);
  const size_t need = (size_t)splits * rows_pad * nb * sizeof(float);
  if (workspace_bytes < need) {
    isa_set_error("renet_proj_wgrad: workspace %zu < %zu", workspace_bytes, need);
    return ISA_ERR_WORKSPACE;
  }
  GemmParams p;
  p.a.nseg = 2; p.a.mn_major = 1; p.a.extent = 8 * n;
  p.a.seg[0] = plain_seg(dgx, 6LL * n, 6 * n, 0);
  p.a.seg[1] = plain_seg(dghn, 2LL * n, 2 * n, 6 * n);
  ISA_CHECK_ARG((6 * n) % 8 == 0, "renet_proj_wgrad: 6 n must be a multiple of 8 (n = %d)", n);
  p.b.nseg = 4; p.b.mn_major = 1; p.b.extent = nb;
  p.b.seg[0] = plain_seg(x, cin, cin, 0);
  p.b.seg[1] = plain_seg(out, 2LL * n, n, cb0);
  p.b.seg[1].shift = -step; p.b.seg[1].pos_div = pos_div; p.b.seg[1].pos_mod = pos_mod; p.b.seg[1].pos_step = -1;
  p.b.seg[2] = plain_seg(out + n, 2LL * n, n, cb1);
  p.b.seg[2].shift = step; p.b.seg[2].pos_div = pos_div; p.b.seg[2].pos_mod = pos_mod; p.b.seg[2].pos_step = 1;
  p.b.seg[3] = plain_seg(nullptr, 0, 1, cone);
  p.b.seg[3].ones = 1;
  p.K = tokens; p.k_per_split = kper;
  p.out = reinterpret_cast<float*>(workspace); p.ldc = nb; p.out_rows = rows_pad; p.out_cols = nb;
  p.bn = pick_bn(nb);
  p.b_packed = nullptr; p.n_iter_total = 0;
  rc = launch_gemm(p, 8LL * n, splits, stream);
  if (rc) return rc;
  const int total = 6 * n * cin + 6 * n * n + 12 * n;
  int blocks = (total + 255) / 256;
  if (blocks > di.num_sms * 8) blocks = di.num_sms * 8;
  renet_wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(workspace), splits, rows_pad, nb, n, cin, cb0, cb1, cone,
                                                        dw_ih, dw_hh, db_ih, db_hh);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
