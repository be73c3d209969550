// ReNet projection GEMMs on tcgen05 with the fp32 -> bf16 hi/lo operand split fused into the producer warps.
//
// The GEMMs around the GRU scan (reference contract /root/reference/code/lib/archs/modules/README.md:225-256; the
// arithmetic is nn.GRU's x W_ih^T and autograd's dG W_ih, dG^T [x | h_prev | 1]) ran as isa_split_bf16x3 (4 B read +
// 6 B written per element) followed by K-concatenated cuBLAS GEMMs over the split copies: ~16 B of HBM traffic per
// operand element against 4 B algorithmic.  Here ONE kernel reads the fp32 operands once:
//   * 8 producer warps load 8-element runs (32 B, a full sector) of both operands straight from global memory, split
//     them in registers into bf16 hi + lo parts (a = hi + lo to ~16 mantissa bits) and store 16-byte rows of the
//     canonical no-swizzle UMMA core matrices into a 2-stage shared-memory ring (conflict free: a quarter warp writes
//     128 contiguous bytes).  The operand may be K-contiguous (rows = M/N index) or MN-contiguous (rows = K index): the
//     store picks the K-major or MN-major core-matrix layout and the instruction descriptor's a_major / b_major bits
//     do the transposition -- no operand is ever transposed or copied in global memory;
//   * one MMA warp issues, per 16-deep k step, a_hi b_hi + a_hi b_lo + a_lo b_hi (M = 128, N <= 256, fp32 accumulation
//     in tensor memory) -- fp32-level accuracy (error ~2^-16 relative) like the rest of the hot path;
//   * mbarriers: full[stage] <- 256 producer arrivals, empty[stage] <- tcgen05.commit; two CTAs per SM (96 KB of shared
//     memory, 256 TMEM columns each) overlap one CTA's epilogue with the other's main loop.
// MN-major operands are concatenations of up to four column segments, each optionally read with a row shift and
// zero-filled where the shifted row leaves its sweep (h_{t-1} of a direction straight from the forward output) or
// all ones (the bias-gradient column), so [dgx | dghn]^T [x | h_prev(0) | h_prev(1) | 1] needs no materialised operand.
// The weight gradient (K = all tokens) is split over blockIdx.z; partials are folded in a fixed order (deterministic).
#include "isa_common.cuh"
#include "isa_tcgen05.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kBNMax = 256;
constexpr int kBK = 32;
constexpr int kStages = 2;
constexpr int kProducerThreads = 256;
constexpr int kThreadsTotal = kProducerThreads + 32;     // + the MMA warp
constexpr int kTileA = kBM * kBK * 2;                    // bytes of one bf16 part of the A stage (8 KB)
constexpr int kTileB = kBNMax * kBK * 2;                 // 16 KB
constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;     // hi + lo of both operands: 48 KB
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
};

__device__ __forceinline__ void load8_kmajor(const Operand& op, long long mn, long long k0, long long K, float (&v)[8]) {
  const Seg& s = op.seg[0];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (mn >= op.extent || k0 >= K) return;
  const float* p = s.ptr + mn * s.ld + k0;
  if (k0 + 8 <= K) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    for (int i = 0; i < 8 && k0 + i < K; ++i) v[i] = __ldg(p + i);
  }
}

__device__ __forceinline__ void load8_mnmajor(const Operand& op, int mn0, long long k, long long K, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (mn0 >= op.extent || k >= K) return;
  int si = 0;
#pragma unroll
  for (int j = 1; j < kMaxSeg; ++j)
    if (j < op.nseg && mn0 >= op.seg[j].v0) si = j;
  const Seg& s = op.seg[si];
  const int c = mn0 - s.v0;
  if (c >= s.cols) return;
  if (s.ones) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c + i < s.cols) ? 1.f : 0.f;
    return;
  }
  if (s.pos_step != 0) {
    const int pos = (int)((k / s.pos_div) % s.pos_mod) + s.pos_step;
    if (pos < 0 || pos >= s.pos_mod) return;
  }
  const float* p = s.ptr + (k + s.shift) * s.ld + c;
  if (c + 8 <= s.cols) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    for (int i = 0; i < 8 && c + i < s.cols; ++i) v[i] = __ldg(p + i);
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

// One operand stage: `groups` 8-wide M/N groups x 4 k-groups of 8.  Chunk q = one 16-byte core-matrix row.
template <int PER_THREAD>
__device__ __forceinline__ void produce(const Operand& op, long long mn_base, int groups, long long k_base, long long K,
                                        unsigned char* hi_tile, unsigned char* lo_tile, int tid) {
  float v[PER_THREAD][8];
  int off[PER_THREAD];
#pragma unroll
  for (int i = 0; i < PER_THREAD; ++i) {
    const int q = tid + i * kProducerThreads;
    off[i] = -1;
    if (op.mn_major) {
      // a warp = 8 k rows x 4 adjacent M/N groups (128 contiguous bytes per global row)
      const int k8 = q & 7, mgl = (q >> 3) & 3, rest = q >> 5, kg = rest & 3, mg = (rest >> 2) * 4 + mgl;
      if (mg < groups) {
        load8_mnmajor(op, (int)mn_base + mg * 8, k_base + kg * 8 + k8, K, v[i]);
        off[i] = mg * kGroupStride + kg * 128 + k8 * 16;
      }
    } else {
      // a warp = one 8-row group, its 4 k-chunks (128 contiguous bytes per global row)
      const int r8 = q & 7, kc = (q >> 3) & 3, rg = q >> 5;
      if (rg < groups) {
        load8_kmajor(op, mn_base + rg * 8 + r8, k_base + kc * 8, K, v[i]);
        off[i] = rg * kGroupStride + kc * 128 + r8 * 16;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PER_THREAD; ++i)
    if (off[i] >= 0) store_split8(hi_tile, lo_tile, off[i], v[i]);
}

__global__ void __launch_bounds__(kThreadsTotal) proj_gemm_kernel(const GemmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw_g[];
  unsigned char* sm = smem_raw_g + ((128u - (smem_u32(smem_raw_g) & 127u)) & 127u);
  __shared__ uint64_t s_full[kStages], s_empty[kStages], s_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long m0 = (long long)blockIdx.y * kBM;
  const int n0 = blockIdx.x * p.bn;
  const long long k_lo = (long long)blockIdx.z * p.k_per_split;
  const long long k_hi = (k_lo + p.k_per_split < p.K) ? k_lo + p.k_per_split : p.K;
  const int n_iter = (int)((k_hi - k_lo + kBK - 1) / kBK);
  int bn = p.b.extent - n0;
  bn = bn > p.bn ? p.bn : ((bn + 15) & ~15);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], kProducerThreads); mbar_init(&s_empty[s], 1); }
    mbar_init(&s_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < kProducerThreads / 32) {
    // ------------------------------------------------------------ producers
    const int a_groups = kBM / 8, b_groups = bn / 8;
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      mbar_wait(&s_empty[s], ph ^ 1u);
      unsigned char* st = sm + s * kStageBytes;
      const long long kb = k_lo + (long long)it * kBK;
      produce<2>(p.a, m0, a_groups, kb, k_hi, st, st + kTileA, tid);
      if (b_groups > 16) produce<4>(p.b, n0, b_groups, kb, k_hi, st + 2 * kTileA, st + 2 * kTileA + kTileB, tid);
      else produce<2>(p.b, n0, b_groups, kb, k_hi, st + 2 * kTileA, st + 2 * kTileA + kTileB, tid);
      fence_proxy_async();
      mbar_arrive(&s_full[s]);
    }
    // ------------------------------------------------------------ epilogue: warps 0..3 own the four TMEM lane quarters
    if (warp < 4) {
      mbar_wait(&s_done, 0);
      tc_fence_after();
      const long long row = m0 + warp * 32 + (tid & 31);
      const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
      float* orow = p.out + ((long long)blockIdx.z * p.out_rows + row) * p.ldc + n0;
      const bool row_ok = row < p.out_rows;
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
        if (row_ok) {
          int lim = p.out_cols - (n0 + c0);
          if (lim > bn - c0) lim = bn - c0;      // the tile's own columns only
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            if (i + 4 <= lim) {
              *reinterpret_cast<uint4*>(orow + c0 + i) = make_uint4(d[i], d[i + 1], d[i + 2], d[i + 3]);
            } else {
              for (int j = 0; j < 4; ++j)
                if (i + j < lim) orow[c0 + i + j] = __uint_as_float(d[i + j]);
            }
          }
        }
      }
      tc_fence_before();
    }
  } else {
    // ------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    const uint32_t el = elect_one();
    const uint32_t idesc = make_idesc(kBM, bn) | ((uint32_t)(p.a.mn_major ? 1 : 0) << 15) | ((uint32_t)(p.b.mn_major ? 1 : 0) << 16);
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      mbar_wait(&s_full[s], ph);
      tc_fence_after();
      const uint32_t base = smem_u32(sm + s * kStageBytes);
      const uint64_t a_hi = make_desc(base, 128, kGroupStride), a_lo = make_desc(base + kTileA, 128, kGroupStride);
      const uint64_t b_hi = make_desc(base + 2 * kTileA, 128, kGroupStride), b_lo = make_desc(base + 2 * kTileA + kTileB, 128, kGroupStride);
#pragma unroll
      for (int kk = 0; kk < kBK / 16; ++kk) {
        const uint64_t ko = (uint64_t)((kk * 256) >> 4);
        umma_bf16_e(el, tmem, a_hi + ko, b_hi + ko, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        umma_bf16_e(el, tmem, a_hi + ko, b_lo + ko, idesc, 1u);
        umma_bf16_e(el, tmem, a_lo + ko, b_hi + ko, idesc, 1u);
      }
      umma_commit_e(el, &s_empty[s]);
    }
    umma_commit_e(el, &s_done);
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
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

int launch_gemm(const GemmParams& p, long long M, int splits, cudaStream_t stream) {
  static thread_local bool attr_set[16] = {false};
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  const size_t smem = (size_t)kStages * kStageBytes + 128;
  if (di.device >= 0 && di.device < 16 && !attr_set[di.device]) {
    ISA_CUDA(cudaFuncSetAttribute(proj_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[di.device] = true;
  }
  dim3 grid((unsigned)((p.b.extent + p.bn - 1) / p.bn), (unsigned)((M + kBM - 1) / kBM), (unsigned)splits);
  proj_gemm_kernel<<<grid, kThreadsTotal, smem, stream>>>(p);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// gx[tokens][n_out] = x[tokens][cin] w[n_out][cin]^T
int isa_renet_proj_fwd(const float* x, const float* w, long long tokens, int cin, int n_out, float* gx, cudaStream_t stream) {
  ISA_CHECK_ARG(x && w && gx, "renet_proj_fwd: null pointer");
  ISA_CHECK_ARG(tokens > 0 && cin > 0 && n_out > 0 && cin % 4 == 0 && n_out % 4 == 0,
                "renet_proj_fwd: tokens=%lld cin=%d n_out=%d (widths must be positive multiples of 4)", tokens, cin, n_out);
  ISA_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(gx), "renet_proj_fwd: pointers must be 16-byte aligned");
  GemmParams p;
  ISA_CHECK_ARG(tokens < 0x7fffffff, "renet_proj_fwd: too many tokens");
  p.a.nseg = 1; p.a.mn_major = 0; p.a.extent = (int)tokens;
  p.a.seg[0] = plain_seg(x, cin, cin, 0);
  p.b.nseg = 1; p.b.mn_major = 0; p.b.extent = n_out;
  p.b.seg[0] = plain_seg(w, cin, cin, 0);
  p.K = cin; p.k_per_split = ((cin + kBK - 1) / kBK) * kBK;
  p.out = gx; p.ldc = n_out; p.out_rows = tokens; p.out_cols = n_out;
  p.bn = pick_bn(n_out);
  return launch_gemm(p, tokens, 1, stream);
}

// dx[tokens][cin] = dg[tokens][n_out] w[n_out][cin]
int isa_renet_proj_dx(const float* dg, const float* w, long long tokens, int n_out, int cin, float* dx, cudaStream_t stream) {
  ISA_CHECK_ARG(dg && w && dx, "renet_proj_dx: null pointer");
  ISA_CHECK_ARG(tokens > 0 && tokens < 0x7fffffff && cin > 0 && n_out > 0 && cin % 4 == 0 && n_out % 4 == 0,
                "renet_proj_dx: tokens=%lld cin=%d n_out=%d (widths must be positive multiples of 4)", tokens, cin, n_out);
  ISA_CHECK_ARG(aligned16(dg) && aligned16(w) && aligned16(dx), "renet_proj_dx: pointers must be 16-byte aligned");
  GemmParams p;
  p.a.nseg = 1; p.a.mn_major = 0; p.a.extent = (int)tokens;
  p.a.seg[0] = plain_seg(dg, n_out, n_out, 0);
  p.b.nseg = 1; p.b.mn_major = 1; p.b.extent = (cin + 7) & ~7;
  p.b.seg[0] = plain_seg(w, cin, cin, 0);
  p.K = n_out; p.k_per_split = ((n_out + kBK - 1) / kBK) * kBK;
  p.out = dx; p.ldc = cin; p.out_rows = tokens; p.out_cols = cin;
  p.bn = pick_bn(p.b.extent);
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
  int s = (int)((2LL * num_sms) / tiles);                  // two CTAs per SM, one wave
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
