// HBM-bound spatial attention primitives of the instance-embedding block
// (/root/reference/code/lib/archs/modules/utils.py):
//   masked softmax over H*W            SpatialAttentionLayer.forward :507-512 (one fg mask per image, times sum(mask))
//                                      HardAttentionLayer.forward   :648-652 (one logit map, K instance masks, NaN -> 0)
//   row reductions / row affine        AttentionLayer (SE) :413-420: global average pool, x * gate; their gradients
//   single-query readout               Decoder.forward :59-69: sigmoid(bmm(q (b,1,C), enc (b,C,HW)))
// Every kernel streams its operands once with 16 B accesses (HW % 4 == 0; a scalar variant covers the rest),
// pixel index fastest, so a warp touches 512 contiguous bytes per load.  Reductions are two-stage and
// deterministic (per-chunk partials in the workspace, combined in a fixed order) -- no float atomics.
#include "isa_common.cuh"
#include <math.h>

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 8192;  // pixels per CTA

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += s_red[i];
  return t;
}

__device__ __forceinline__ void online_merge(float& m, float& l, float m2, float l2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { l = 0.f; m = mn; return; }
  l = l * __expf(m - mn) + l2 * __expf(m2 - mn);
  m = mn;
}

template <typename MT>
__device__ __forceinline__ bool keep_of(MT v) { return v != (MT)0; }

// ------------------------------------------------------------------------------------ masked softmax
// pass A: per (row, chunk) running max / sum over kept pixels + the keep bitmask (32 pixels per word)
template <typename MT>
__global__ void __launch_bounds__(kThreads) msm_stats_kernel(const float* __restrict__ x, const MT* __restrict__ mask, int K, int HW,
                                                             int n_chunks, float2* __restrict__ part, uint32_t* __restrict__ bits) {
  __shared__ float s_m[kThreads / 32], s_l[kThreads / 32];
  const int row = blockIdx.y, chunk = blockIdx.x;
  const int b = row / K;
  const float* xr = x + (size_t)b * HW;
  const MT* mr = mask + (size_t)row * HW;
  const int words_per_row = (HW + 31) >> 5;
  uint32_t* br = bits + (size_t)row * words_per_row;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  float m = -INFINITY, l = 0.f;
  // one warp iteration = 32 pixels = one bitmask word
  for (int base = p0 + (threadIdx.x >> 5) * 32; base < p1; base += kThreads) {
    const int p = base + (threadIdx.x & 31);
    bool keep = false;
    float v = 0.f;
    if (p < p1) {
      keep = keep_of(mr[p]);
      v = xr[p];
    }
    const uint32_t word = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) br[base >> 5] = word;
    if (keep) {
      if (v > m) { l = l * __expf(m - v) + 1.f; m = v; }
      else l += __expf(v - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    online_merge(m, l, m2, l2);
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_l[threadIdx.x >> 5] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kThreads / 32; ++i) online_merge(m, l, s_m[i], s_l[i]);
    part[(size_t)row * n_chunks + chunk] = make_float2(m, l);
  }
}

// pass B: combine the row's partials (fixed order), write y = keep ? exp(x - M) / L * scale : 0
__global__ void __launch_bounds__(kThreads) msm_write_kernel(const float* __restrict__ x, const uint32_t* __restrict__ bits, int K, int HW,
                                                             int n_chunks, const float2* __restrict__ part, const float* __restrict__ scale,
                                                             int nan_to_zero, float* __restrict__ y, float2* __restrict__ stats) {
  __shared__ float s_M, s_inv;
  const int row = blockIdx.y, chunk = blockIdx.x;
  const int b = row / K;
  if (threadIdx.x == 0) {
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_chunks; ++i) {
      const float2 q = part[(size_t)row * n_chunks + i];
      online_merge(m, l, q.x, q.y);
    }
    s_M = m;
    // empty mask: the reference's softmax over all -inf is NaN on the whole row (utils.py:510, :649)
    s_inv = l > 0.f ? (scale ? scale[row] : 1.f) / l : (nan_to_zero ? 0.f : NAN);
    if (chunk == 0 && stats) stats[row] = make_float2(m, l);
  }
  __syncthreads();
  const float M = s_M, inv = s_inv;
  const bool empty = !(inv == inv) || (M == -INFINITY);
  const float* xr = x + (size_t)b * HW;
  const int words_per_row = (HW + 31) >> 5;
  const uint32_t* br = bits + (size_t)row * words_per_row;
  float* yr = y + (size_t)row * HW;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
    float out;
    if (empty) out = inv;  // NaN or 0 everywhere
    else {
      const bool keep = (br[p >> 5] >> (p & 31)) & 1u;
      out = keep ? __expf(xr[p] - M) * inv : 0.f;
    }
    yr[p] = out;
  }
}

// backward pass A: D[row] partials = sum_p y dy   (only rows with a non-empty mask matter)
__global__ void __launch_bounds__(kThreads) rowdot_part_kernel(const float* __restrict__ a, const float* __restrict__ bvec, int b_rows_div, int HW,
                                                               int n_chunks, float* __restrict__ part) {
  __shared__ float s_red[kThreads / 32];
  const int row = blockIdx.y, chunk = blockIdx.x;
  const float* ar = a + (size_t)row * HW;
  const float* br = bvec ? bvec + (size_t)(row / b_rows_div) * HW : nullptr;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  float acc = 0.f;
  const bool vec = (HW & 3) == 0;
  if (vec) {
    for (int p = p0 + threadIdx.x * 4; p < p1; p += kThreads * 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(ar + p));
      if (br) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(br + p));
        acc += u.x * w.x + u.y * w.y + u.z * w.z + u.w * w.w;
      } else acc += (u.x + u.y) + (u.z + u.w);
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += kThreads) acc += br ? ar[p] * br[p] : ar[p];
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) part[(size_t)row * n_chunks + chunk] = acc;
}

__global__ void rowdot_final_kernel(const float* __restrict__ part, int rows, int n_chunks, float* __restrict__ out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float acc = 0.f;
  for (int i = 0; i < n_chunks; ++i) acc += part[(size_t)row * n_chunks + i];
  out[row] = acc;
}

// backward pass B: dx[b][p] = sum_k y[b][k][p] * (dy[b][k][p] - D[b][k] / scale[b][k]); rows with an empty mask are skipped
__global__ void __launch_bounds__(kThreads) msm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ D,
                                                           const float2* __restrict__ stats, const float* __restrict__ scale, int K, int HW,
                                                           float* __restrict__ dx) {
  extern __shared__ float s_c[];  // [K] D_k / scale_k, or NaN marker for skipped rows
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += kThreads) {
    const int row = b * K + k;
    const float l = stats[row].y;
    const float sc = scale ? scale[row] : 1.f;
    s_c[k] = (l > 0.f && sc != 0.f) ? D[row] / sc : (l > 0.f ? 0.f : NAN);
  }
  __syncthreads();
  const int p0 = blockIdx.x * kChunk, p1 = min(HW, p0 + kChunk);
  const float* yb = y + (size_t)b * K * HW;
  const float* db = dy + (size_t)b * K * HW;
  float* dxb = dx + (size_t)b * HW;
  if ((HW & 3) == 0) {
    for (int p = p0 + threadIdx.x * 4; p < p1; p += kThreads * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < K; ++k) {
        const float c = s_c[k];
        if (c != c) continue;
        const float4 u = __ldg(reinterpret_cast<const float4*>(yb + (size_t)k * HW + p));
        const float4 g = __ldg(reinterpret_cast<const float4*>(db + (size_t)k * HW + p));
        acc.x += u.x * (g.x - c); acc.y += u.y * (g.y - c); acc.z += u.z * (g.z - c); acc.w += u.w * (g.w - c);
      }
      *reinterpret_cast<float4*>(dxb + p) = acc;
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const float c = s_c[k];
        if (c != c) continue;
        acc += yb[(size_t)k * HW + p] * (db[(size_t)k * HW + p] - c);
      }
      dxb[p] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------ row affine: y = x * g[row] + c[row]
__global__ void __launch_bounds__(kThreads) row_affine_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ c,
                                                              int HW, float* __restrict__ y) {
  const int row = blockIdx.y;
  const float gv = g[row], cv = c ? c[row] : 0.f;
  const float* xr = x + (size_t)row * HW;
  float* yr = y + (size_t)row * HW;
  const int p0 = blockIdx.x * kChunk, p1 = min(HW, p0 + kChunk);
  if ((HW & 3) == 0) {
    for (int p = p0 + threadIdx.x * 4; p < p1; p += kThreads * 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(xr + p));
      *reinterpret_cast<float4*>(yr + p) = make_float4(fmaf(u.x, gv, cv), fmaf(u.y, gv, cv), fmaf(u.z, gv, cv), fmaf(u.w, gv, cv));
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += kThreads) yr[p] = fmaf(xr[p], gv, cv);
  }
}

// ------------------------------------------------------------------------------------ single-query readout
// out[b][p] = sigmoid(sum_c q[b][c] enc[b][c][p])
__global__ void __launch_bounds__(kThreads) readout_fwd_kernel(const float* __restrict__ q, const float* __restrict__ enc, int C, int HW,
                                                               float* __restrict__ out) {
  extern __shared__ float s_q[];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kThreads) s_q[c] = q[(size_t)b * C + c];
  __syncthreads();
  const float* eb = enc + (size_t)b * C * HW;
  const int p0 = blockIdx.x * kChunk, p1 = min(HW, p0 + kChunk);
  if ((HW & 3) == 0) {
    for (int p = p0 + threadIdx.x * 4; p < p1; p += kThreads * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < C; ++c) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(eb + (size_t)c * HW + p));
        const float w = s_q[c];
        acc.x = fmaf(w, u.x, acc.x); acc.y = fmaf(w, u.y, acc.y); acc.z = fmaf(w, u.z, acc.z); acc.w = fmaf(w, u.w, acc.w);
      }
      *reinterpret_cast<float4*>(out + (size_t)b * HW + p) =
          make_float4(1.f / (1.f + expf(-acc.x)), 1.f / (1.f + expf(-acc.y)), 1.f / (1.f + expf(-acc.z)), 1.f / (1.f + expf(-acc.w)));
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(s_q[c], eb[(size_t)c * HW + p], acc);
      out[(size_t)b * HW + p] = 1.f / (1.f + expf(-acc));
    }
  }
}

// dz = dout * o * (1 - o);  denc[b][c][p] = q[b][c] * dz[b][p]   (dq = row-dot of enc with dz, separate call)
__global__ void __launch_bounds__(kThreads) readout_bwd_kernel(const float* __restrict__ q, const float* __restrict__ out, const float* __restrict__ dout,
                                                               int C, int HW, float* __restrict__ dz, float* __restrict__ denc) {
  extern __shared__ float s_q[];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kThreads) s_q[c] = q[(size_t)b * C + c];
  __syncthreads();
  const int p0 = blockIdx.x * kChunk, p1 = min(HW, p0 + kChunk);
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const float o = out[(size_t)b * HW + p];
    const float g = dout[(size_t)b * HW + p] * o * (1.f - o);
    dz[(size_t)b * HW + p] = g;
    if (denc)
      for (int c = 0; c < C; ++c) denc[((size_t)b * C + c) * HW + p] = s_q[c] * g;
  }
}

// ------------------------------------------------------------------ masked batch norm (utils.py:529-591 maskBN, training mode)
//   mm[b]   = sum over ALL channels and pixels of the mask + 1            (utils.py:574)
//   mean[c] = (1/B) sum_b (sum_p x m) / mm[b]          var[c] = (1/B) sum_b (sum_p (x - mean[c])^2 m) / mm[b]
//   y       = (x - mean) / sqrt(var + eps) * weight + bias                (every pixel, masked or not)
// and the gradient flows through mean AND var like the reference's autograd graph does.  All sums are two-stage and
// folded in a fixed order (per-chunk block sums -> loops over chunks and images): deterministic.
// msum [B][MC] = per-row mask sums (MC = 1: one mask plane shared by the channels, MC = C: one per channel).
__device__ __forceinline__ float mbn_mask_mean(const float* __restrict__ msum, int b, int C, int MC) {
  float t = 0.f;
  if (MC == 1) t = (float)C * msum[b];
  else for (int c = 0; c < C; ++c) t += msum[(size_t)b * MC + c];
  return t + 1.f;
}
// fold of a [B][C][nc] partial array over chunks and images for channel c, each image weighted by 1 / (B mm[b])
__device__ __forceinline__ float mbn_fold(const float* __restrict__ part, const float* __restrict__ msum, int B, int C, int MC, int nc, int c,
                                          bool weighted) {
  float tot = 0.f;
  for (int b = 0; b < B; ++b) {
    float r = 0.f;
    const float* pr = part + ((size_t)b * C + c) * nc;
    for (int i = 0; i < nc; ++i) r += pr[i];
    tot += weighted ? r / (mbn_mask_mean(msum, b, C, MC) * (float)B) : r;
  }
  return tot;
}

// part2[row][chunk] = sum_p (x - mean[c])^2 m
__global__ void __launch_bounds__(kThreads) mbn_var_part_kernel(const float* __restrict__ x, const float* __restrict__ mask, int MC,
                                                                const float* __restrict__ part1, const float* __restrict__ msum, int B, int C,
                                                                int HW, int nc, float* __restrict__ part2) {
  __shared__ float s_red[kThreads / 32];
  __shared__ float s_mean;
  const int row = blockIdx.y, chunk = blockIdx.x, b = row / C, c = row - b * C;
  if (threadIdx.x == 0) s_mean = mbn_fold(part1, msum, B, C, MC, nc, c, true);
  __syncthreads();
  const float mean = s_mean;
  const float* xr = x + (size_t)row * HW;
  const float* mr = mask + ((size_t)b * MC + (MC == 1 ? 0 : c)) * HW;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  float acc = 0.f;
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const float d = xr[p] - mean;
    acc = fmaf(d * d, mr[p], acc);
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) part2[(size_t)row * nc + chunk] = acc;
}

// y = (x - mean) * rstd * w + bias; the first CTA of every channel also records (mean, var)
__global__ void __launch_bounds__(kThreads) mbn_norm_kernel(const float* __restrict__ x, const float* __restrict__ part1, const float* __restrict__ part2,
                                                            const float* __restrict__ msum, const float* __restrict__ weight,
                                                            const float* __restrict__ bias, float eps, int B, int C, int MC, int HW, int nc,
                                                            float* __restrict__ y, float* __restrict__ stats) {
  __shared__ float s_mv[2];
  const int row = blockIdx.y, chunk = blockIdx.x, b = row / C, c = row - b * C;
  if (threadIdx.x == 0) {
    s_mv[0] = mbn_fold(part1, msum, B, C, MC, nc, c, true);
    s_mv[1] = mbn_fold(part2, msum, B, C, MC, nc, c, true);
    if (b == 0 && chunk == 0) { stats[c] = s_mv[0]; stats[C + c] = s_mv[1]; }
  }
  __syncthreads();
  const float mean = s_mv[0], rstd = 1.f / sqrtf(s_mv[1] + eps);
  const float g = (weight ? weight[c] : 1.f) * rstd, sh = bias ? bias[c] : 0.f;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  const float* xr = x + (size_t)row * HW;
  float* yr = y + (size_t)row * HW;
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) yr[p] = fmaf(xr[p] - mean, g, sh);
}

// g1[row][chunk] = sum_p dy, g2[row][chunk] = sum_p dy (x - mean[c])
__global__ void __launch_bounds__(kThreads) mbn_bwd_part_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                                                int C, int HW, int nc, float* __restrict__ g1, float* __restrict__ g2) {
  __shared__ float s_red[kThreads / 32];
  const int row = blockIdx.y, chunk = blockIdx.x, c = row % C;
  const float mean = stats[c];
  const float* xr = x + (size_t)row * HW;
  const float* gr = dy + (size_t)row * HW;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  float a1 = 0.f, a2 = 0.f;
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const float g = gr[p];
    a1 += g;
    a2 = fmaf(g, xr[p] - mean, a2);
  }
  a1 = block_sum(a1, s_red);
  __syncthreads();
  a2 = block_sum(a2, s_red);
  if (threadIdx.x == 0) { g1[(size_t)row * nc + chunk] = a1; g2[(size_t)row * nc + chunk] = a2; }
}

// dx = dy r w + m / (B mm[b]) * (dL/dmean + 2 dL/dvar (x - mean));  dweight = r sum dy (x - mean), dbias = sum dy
__global__ void __launch_bounds__(kThreads) mbn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ mask, int MC,
                                                              const float* __restrict__ dy, const float* __restrict__ stats,
                                                              const float* __restrict__ msum, const float* __restrict__ weight, float eps,
                                                              const float* __restrict__ g1, const float* __restrict__ g2, int B, int C, int HW, int nc,
                                                              float* __restrict__ dx, float* __restrict__ dweight, float* __restrict__ dbias) {
  __shared__ float s_v[4];
  const int row = blockIdx.y, chunk = blockIdx.x, b = row / C, c = row - b * C;
  if (threadIdx.x == 0) {
    const float mean = stats[c], rstd = 1.f / sqrtf(stats[C + c] + eps), w = weight ? weight[c] : 1.f;
    const float G1 = mbn_fold(g1, msum, B, C, MC, nc, c, false), G2 = mbn_fold(g2, msum, B, C, MC, nc, c, false);
    // A = sum_b (sum_p m) / (B mm[b]): the weights of the masked mean do not add up to one (the "+ 1", and mm counts
    // every channel), so d var / d mean = -2 mean_of_(x - mean) = -2 mean (1 - A) does not vanish
    float A = 0.f;
    for (int bb = 0; bb < B; ++bb) A += msum[(size_t)bb * MC + (MC == 1 ? 0 : c)] / (mbn_mask_mean(msum, bb, C, MC) * (float)B);
    const float dvar = -0.5f * w * rstd * rstd * rstd * G2;
    const float dmean = -rstd * w * G1 + dvar * (-2.f * mean * (1.f - A));
    s_v[0] = rstd * w; s_v[1] = dmean; s_v[2] = 2.f * dvar; s_v[3] = mean;
    if (b == 0 && chunk == 0) {
      if (dweight) dweight[c] = rstd * G2;
      if (dbias) dbias[c] = G1;
    }
  }
  __syncthreads();
  const float rw = s_v[0], dmean = s_v[1], dvar2 = s_v[2], mean = s_v[3];
  const float ab = 1.f / (mbn_mask_mean(msum, b, C, MC) * (float)B);
  const float* xr = x + (size_t)row * HW;
  const float* gr = dy + (size_t)row * HW;
  const float* mr = mask + ((size_t)b * MC + (MC == 1 ? 0 : c)) * HW;
  float* dr = dx + (size_t)row * HW;
  const int p0 = chunk * kChunk, p1 = min(HW, p0 + kChunk);
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) dr[p] = fmaf(gr[p], rw, ab * mr[p] * fmaf(dvar2, xr[p] - mean, dmean));
}

int chunks_of(int HW) { return (HW + kChunk - 1) / kChunk; }

}  // namespace

extern "C" {

size_t isa_masked_softmax_hw_workspace_bytes(int B, int K, int HW) {
  const size_t rows = (size_t)B * K;
  return isa_align_up(rows * chunks_of(HW) * sizeof(float2), 256) + isa_align_up(rows * ((HW + 31) / 32) * sizeof(uint32_t), 256);
}

int isa_masked_softmax_hw_fwd(const float* x, const void* mask, int mask_kind, int B, int K, int HW, const float* scale, int nan_to_zero,
                              float* y, float* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(x && mask && y && workspace, "masked_softmax_hw_fwd: null pointer");
  ISA_CHECK_ARG(B > 0 && K > 0 && HW > 0 && (long long)B * K <= 65535, "masked_softmax_hw_fwd: bad sizes (B=%d K=%d HW=%d)", B, K, HW);
  ISA_CHECK_ARG(mask_kind == 0 || mask_kind == 1, "masked_softmax_hw_fwd: mask_kind must be 0 (u8) or 1 (f32)");
  ISA_CHECK_ARG(workspace_bytes >= isa_masked_softmax_hw_workspace_bytes(B, K, HW), "masked_softmax_hw_fwd: workspace too small");
  const int rows = B * K, nc = chunks_of(HW);
  float2* part = reinterpret_cast<float2*>(workspace);
  uint32_t* bits = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + isa_align_up((size_t)rows * nc * sizeof(float2), 256));
  dim3 grid(nc, rows);
  if (mask_kind == 0)
    msm_stats_kernel<uint8_t><<<grid, kThreads, 0, stream>>>(x, reinterpret_cast<const uint8_t*>(mask), K, HW, nc, part, bits);
  else
    msm_stats_kernel<float><<<grid, kThreads, 0, stream>>>(x, reinterpret_cast<const float*>(mask), K, HW, nc, part, bits);
  ISA_CUDA(cudaGetLastError());
  msm_write_kernel<<<grid, kThreads, 0, stream>>>(x, bits, K, HW, nc, part, scale, nan_to_zero, y, reinterpret_cast<float2*>(stats));
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_masked_softmax_hw_bwd(const float* y, const float* dy, const float* stats, const float* scale, int B, int K, int HW, float* dx,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(y && dy && stats && dx && workspace, "masked_softmax_hw_bwd: null pointer");
  ISA_CHECK_ARG(B > 0 && K > 0 && HW > 0 && (long long)B * K <= 65535 && K <= 8192, "masked_softmax_hw_bwd: bad sizes");
  const int rows = B * K, nc = chunks_of(HW);
  ISA_CHECK_ARG(workspace_bytes >= isa_masked_softmax_hw_workspace_bytes(B, K, HW), "masked_softmax_hw_bwd: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  float* D = part + (size_t)rows * nc;
  rowdot_part_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(y, dy, 1, HW, nc, part);
  ISA_CUDA(cudaGetLastError());
  rowdot_final_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(part, rows, nc, D);
  ISA_CUDA(cudaGetLastError());
  msm_bwd_kernel<<<dim3(nc, B), kThreads, K * sizeof(float), stream>>>(y, dy, D, reinterpret_cast<const float2*>(stats), scale, K, HW, dx);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

size_t isa_row_dot_workspace_bytes(int rows, int HW) { return isa_align_up((size_t)rows * chunks_of(HW) * sizeof(float), 256); }

int isa_row_dot(const float* a, const float* b, int rows, int HW, int b_rows_div, float* out, void* workspace, size_t workspace_bytes,
                cudaStream_t stream) {
  ISA_CHECK_ARG(a && out && workspace, "row_dot: null pointer");
  ISA_CHECK_ARG(rows > 0 && rows <= 65535 && HW > 0 && b_rows_div > 0, "row_dot: bad sizes (rows=%d HW=%d)", rows, HW);
  ISA_CHECK_ARG(workspace_bytes >= isa_row_dot_workspace_bytes(rows, HW), "row_dot: workspace too small");
  const int nc = chunks_of(HW);
  float* part = reinterpret_cast<float*>(workspace);
  rowdot_part_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(a, b, b_rows_div, HW, nc, part);
  ISA_CUDA(cudaGetLastError());
  rowdot_final_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(part, rows, nc, out);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_row_affine(const float* x, const float* g, const float* c, int rows, int HW, float* y, cudaStream_t stream) {
  ISA_CHECK_ARG(x && g && y, "row_affine: null pointer");
  ISA_CHECK_ARG(rows > 0 && rows <= 65535 && HW > 0, "row_affine: bad sizes (rows=%d HW=%d)", rows, HW);
  row_affine_kernel<<<dim3(chunks_of(HW), rows), kThreads, 0, stream>>>(x, g, c, HW, y);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

size_t isa_mask_bn_workspace_bytes(int B, int C, int HW) {
  return 4 * isa_align_up((size_t)B * C * chunks_of(HW) * sizeof(float), 256);
}

int isa_mask_bn_fwd(const float* x, const float* mask, int mask_channels, int B, int C, int HW, const float* weight, const float* bias, float eps,
                    float* y, float* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(x && mask && y && stats && workspace, "mask_bn_fwd: null pointer");
  ISA_CHECK_ARG(B > 0 && C > 0 && HW > 0 && (long long)B * C <= 65535, "mask_bn_fwd: bad sizes (B=%d C=%d HW=%d)", B, C, HW);
  ISA_CHECK_ARG(mask_channels == 1 || mask_channels == C, "mask_bn_fwd: the mask has 1 or C channels (got %d)", mask_channels);
  ISA_CHECK_ARG(workspace_bytes >= isa_mask_bn_workspace_bytes(B, C, HW), "mask_bn_fwd: workspace too small");
  const int rows = B * C, nc = chunks_of(HW), MC = mask_channels;
  const size_t stride = isa_align_up((size_t)rows * nc * sizeof(float), 256) / sizeof(float);
  float* part1 = reinterpret_cast<float*>(workspace);
  float* part2 = part1 + stride;
  float* pm = part2 + stride;
  float* msum = stats + 2 * C;                                           // [B][MC], kept for the backward call
  rowdot_part_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(x, mask, MC == 1 ? C : 1, HW, nc, part1);
  ISA_CUDA(cudaGetLastError());
  rowdot_part_kernel<<<dim3(nc, B * MC), kThreads, 0, stream>>>(mask, nullptr, 1, HW, nc, pm);
  ISA_CUDA(cudaGetLastError());
  rowdot_final_kernel<<<(B * MC + 127) / 128, 128, 0, stream>>>(pm, B * MC, nc, msum);
  ISA_CUDA(cudaGetLastError());
  mbn_var_part_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(x, mask, MC, part1, msum, B, C, HW, nc, part2);
  ISA_CUDA(cudaGetLastError());
  mbn_norm_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(x, part1, part2, msum, weight, bias, eps, B, C, MC, HW, nc, y, stats);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_mask_bn_bwd(const float* x, const float* mask, int mask_channels, const float* dy, const float* stats, const float* weight, float eps,
                    int B, int C, int HW, float* dx, float* dweight, float* dbias, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(x && mask && dy && stats && dx && workspace, "mask_bn_bwd: null pointer");
  ISA_CHECK_ARG(B > 0 && C > 0 && HW > 0 && (long long)B * C <= 65535, "mask_bn_bwd: bad sizes (B=%d C=%d HW=%d)", B, C, HW);
  ISA_CHECK_ARG(mask_channels == 1 || mask_channels == C, "mask_bn_bwd: the mask has 1 or C channels (got %d)", mask_channels);
  ISA_CHECK_ARG(workspace_bytes >= isa_mask_bn_workspace_bytes(B, C, HW), "mask_bn_bwd: workspace too small");
  const int rows = B * C, nc = chunks_of(HW), MC = mask_channels;
  const size_t stride = isa_align_up((size_t)rows * nc * sizeof(float), 256) / sizeof(float);
  float* g1 = reinterpret_cast<float*>(workspace);
  float* g2 = g1 + stride;
  const float* msum = stats + 2 * C;
  mbn_bwd_part_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(x, dy, stats, C, HW, nc, g1, g2);
  ISA_CUDA(cudaGetLastError());
  mbn_bwd_dx_kernel<<<dim3(nc, rows), kThreads, 0, stream>>>(x, mask, MC, dy, stats, msum, weight, eps, g1, g2, B, C, HW, nc, dx, dweight, dbias);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_readout_fwd(const float* q, const float* enc, int B, int C, int HW, float* out, cudaStream_t stream) {
  ISA_CHECK_ARG(q && enc && out, "readout_fwd: null pointer");
  ISA_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && C <= 4096 && HW > 0, "readout_fwd: bad sizes");
  readout_fwd_kernel<<<dim3(chunks_of(HW), B), kThreads, C * sizeof(float), stream>>>(q, enc, C, HW, out);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_readout_bwd(const float* q, const float* out, const float* dout, int B, int C, int HW, float* dz, float* denc, cudaStream_t stream) {
  ISA_CHECK_ARG(q && out && dout && dz, "readout_bwd: null pointer");
  ISA_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && C <= 4096 && HW > 0, "readout_bwd: bad sizes");
  readout_bwd_kernel<<<dim3(chunks_of(HW), B), kThreads, C * sizeof(float), stream>>>(q, out, dout, C, HW, dz, denc);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
