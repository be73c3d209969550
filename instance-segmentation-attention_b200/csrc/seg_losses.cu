// Fused semantic-head losses: cross entropy + Dice over the (bs, n_classes, H, W) logits in ONE pass
// (forward) and ONE pass (backward).  Replaces, on the training step next to the discriminative loss,
//   /root/reference/code/lib/losses/dice.py:10-51 (dice_coefficient), :54-89 (dice_loss) and the
//   CrossEntropyLoss call of /root/reference/code/lib/model.py:255-263
// which in PyTorch are ~20 elementwise / reduce launches over the same 9 bytes per pixel.
//
// HBM-bound: forward reads 4*nc + 1 bytes per pixel, backward reads the same and writes 4*nc.
// Deterministic: per-CTA partial sums (double) are folded in a fixed order by the last CTA to finish.
#include "isa_common.cuh"

namespace {

constexpr int SL_MAX_NC = 8;
constexpr int SL_THREADS = 256;
constexpr int SL_PIX_PER_THREAD = 8;
constexpr int SL_PIX_PER_CTA = SL_THREADS * SL_PIX_PER_THREAD;

// workspace layout (doubles unless stated)
//   [0]                      ticket (unsigned, first 8 bytes)
//   sums   [bs][nc][3]       N = sum p t, A = sum f(p), B = sum f(t)
//   ce     [2]               sum w[t] * nll, sum w[t]
//   coef   [bs][nc][2]       backward coefficients alpha, beta (see below)
//   part   [n_cta][3*nc+2]   per-CTA partials
struct SegLayout {
  size_t off_sums, off_ce, off_coef, off_part, total;
  int chunks;
};

__host__ __device__ inline SegLayout seg_layout(int bs, int nc, long long HW) {
  SegLayout L;
  L.chunks = (int)((HW + SL_PIX_PER_CTA - 1) / SL_PIX_PER_CTA);
  size_t o = 16;
  L.off_sums = o; o += sizeof(double) * (size_t)bs * nc * 3;
  L.off_ce = o;   o += sizeof(double) * 2;
  L.off_coef = o; o += sizeof(double) * (size_t)bs * nc * 2;
  L.off_part = o; o += sizeof(double) * (size_t)bs * L.chunks * (3 * nc + 2);
  L.total = o;
  return L;
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < SL_THREADS / 32) ? sh[l] : 0.0;
    r = warp_sum_d(r);
  }
  return r;  // valid in warp 0
}

// out[0] = CE, out[1] = Dice loss
__global__ void __launch_bounds__(SL_THREADS)
seg_losses_fwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ labels,
                      const float* __restrict__ class_w, int bs, int nc, long long HW, int time2, float smooth,
                      int optimize_bg, unsigned char* __restrict__ ws, float* __restrict__ out) {
  const SegLayout L = seg_layout(bs, nc, HW);
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* z = logits + (size_t)b * nc * HW;
  const unsigned char* lab = labels + (size_t)b * HW;
  float accN[SL_MAX_NC], accA[SL_MAX_NC], accB[SL_MAX_NC];
#pragma unroll
  for (int c = 0; c < SL_MAX_NC; ++c) accN[c] = accA[c] = accB[c] = 0.f;
  float ce_num = 0.f, ce_den = 0.f;
  const long long base = (long long)chunk * SL_PIX_PER_CTA;
#pragma unroll 2
  for (int i = 0; i < SL_PIX_PER_THREAD; ++i) {
    const long long p = base + (long long)i * SL_THREADS + threadIdx.x;
    if (p >= HW) break;
    float v[SL_MAX_NC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) { v[c] = __ldg(z + (size_t)c * HW + p); m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) { v[c] = __expf(v[c] - m); s += v[c]; }
    const float inv = 1.f / s;
    const int t = lab[p];
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) {
        const float pc = v[c] * inv;
        const float tc = (c == t) ? 1.f : 0.f;
        accN[c] += pc * tc;
        accA[c] += time2 ? pc * pc : pc;
        accB[c] += tc;
        if (c == t) {
          const float w = class_w ? class_w[c] : 1.f;
          ce_num += w * (-__logf(pc));
          ce_den += w;
        }
      }
  }
  __shared__ double sh[SL_THREADS / 32];
  double* part = reinterpret_cast<double*>(ws + L.off_part) + ((size_t)b * L.chunks + chunk) * (3 * nc + 2);
  for (int c = 0; c < nc; ++c) {
    double r;
    r = block_sum_d((double)accN[c], sh); if (threadIdx.x == 0) part[3 * c + 0] = r;
    r = block_sum_d((double)accA[c], sh); if (threadIdx.x == 0) part[3 * c + 1] = r;
    r = block_sum_d((double)accB[c], sh); if (threadIdx.x == 0) part[3 * c + 2] = r;
  }
  {
    double r;
    r = block_sum_d((double)ce_num, sh); if (threadIdx.x == 0) part[3 * nc + 0] = r;
    r = block_sum_d((double)ce_den, sh); if (threadIdx.x == 0) part[3 * nc + 1] = r;
  }
  // last CTA folds the partials in a fixed order
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(ws);
    const unsigned total = gridDim.x * gridDim.y;
    is_last = (atomicAdd(ticket, 1u) == total - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double* sums = reinterpret_cast<double*>(ws + L.off_sums);
  double* ce = reinterpret_cast<double*>(ws + L.off_ce);
  double* coef = reinterpret_cast<double*>(ws + L.off_coef);
  const double* allp = reinterpret_cast<const double*>(ws + L.off_part);
  const int per = 3 * nc + 2;
  // thread (b', j) sums column j of image b' over the chunks, in chunk order
  for (int idx = threadIdx.x; idx < bs * 3 * nc; idx += SL_THREADS) {
    const int bb = idx / (3 * nc), j = idx % (3 * nc);
    double a = 0.0;
    for (int ch = 0; ch < L.chunks; ++ch) a += allp[((size_t)bb * L.chunks + ch) * per + j];
    sums[(size_t)bb * 3 * nc + j] = a;
  }
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int bb = 0; bb < bs; ++bb)
      for (int ch = 0; ch < L.chunks; ++ch) a += allp[((size_t)bb * L.chunks + ch) * per + 3 * nc + threadIdx.x];
    ce[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int c0 = optimize_bg ? 0 : 1;
    const int nsel = nc - c0;
    double wsum = 0.0;
    for (int c = c0; c < nc; ++c) wsum += class_w ? (double)class_w[c] : 1.0;
    double dice_loss = 0.0;
    for (int bb = 0; bb < bs; ++bb) {
      double acc = 0.0;
      for (int c = 0; c < nc; ++c) {
        const double N = sums[((size_t)bb * nc + c) * 3 + 0], A = sums[((size_t)bb * nc + c) * 3 + 1],
                     B = sums[((size_t)bb * nc + c) * 3 + 2];
        const double den = A + B + (double)smooth;
        const double D = (2.0 * N + (double)smooth) / den;
        double wn = 0.0;
        if (c >= c0 && nsel > 0) wn = class_w ? (double)nsel * (double)class_w[c] / wsum : 1.0;
        if (c >= c0) acc += wn * D;
        // dL/dD = -wn / (bs * nsel);  dD/dp = [2 t den - (2N+s) f'(p)] / den^2
        const double dLdD = (c >= c0 && nsel > 0) ? -wn / ((double)bs * (double)nsel) : 0.0;
        coef[((size_t)bb * nc + c) * 2 + 0] = dLdD * 2.0 / den;                               // alpha: * t
        coef[((size_t)bb * nc + c) * 2 + 1] = -dLdD * (2.0 * N + (double)smooth) / (den * den);  // beta: * f'(p)
      }
      dice_loss += 1.0 - (nsel > 0 ? acc / (double)nsel : 0.0);
    }
    out[0] = (float)(ce[0] / ce[1]);
    out[1] = (float)(dice_loss / (double)bs);
    *reinterpret_cast<unsigned*>(ws) = 0u;  // ticket reusable (graph replay)
  }
}

__global__ void __launch_bounds__(SL_THREADS)
seg_losses_bwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ labels,
                      const float* __restrict__ class_w, int bs, int nc, long long HW, int time2,
                      const unsigned char* __restrict__ ws, const float* __restrict__ g_ce,
                      const float* __restrict__ g_dice, float* __restrict__ grad) {
  const SegLayout L = seg_layout(bs, nc, HW);
  const int b = blockIdx.y;
  const float* z = logits + (size_t)b * nc * HW;
  float* gz = grad + (size_t)b * nc * HW;
  const unsigned char* lab = labels + (size_t)b * HW;
  const double* ce = reinterpret_cast<const double*>(ws + L.off_ce);
  const double* coef = reinterpret_cast<const double*>(ws + L.off_coef) + (size_t)b * nc * 2;
  const float gce = g_ce ? g_ce[0] : 0.f, gd = g_dice ? g_dice[0] : 0.f;
  const float inv_ce_den = (float)(1.0 / ce[1]);
  float alpha[SL_MAX_NC], beta[SL_MAX_NC], wcl[SL_MAX_NC];
#pragma unroll
  for (int c = 0; c < SL_MAX_NC; ++c) {
    alpha[c] = c < nc ? (float)coef[2 * c] * gd : 0.f;
    beta[c] = c < nc ? (float)coef[2 * c + 1] * gd : 0.f;
    wcl[c] = c < nc ? (class_w ? class_w[c] : 1.f) : 0.f;
  }
  const long long base = (long long)blockIdx.x * SL_PIX_PER_CTA;
#pragma unroll 2
  for (int i = 0; i < SL_PIX_PER_THREAD; ++i) {
    const long long p = base + (long long)i * SL_THREADS + threadIdx.x;
    if (p >= HW) break;
    float v[SL_MAX_NC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) { v[c] = __ldg(z + (size_t)c * HW + p); m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) { v[c] = __expf(v[c] - m); s += v[c]; }
    const float inv = 1.f / s;
    const int t = lab[p];
    float g[SL_MAX_NC];
    float dotgp = 0.f;
    float wt = 0.f;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) {
        v[c] *= inv;
        const float fp = time2 ? 2.f * v[c] : 1.f;
        g[c] = (c == t ? alpha[c] : 0.f) + beta[c] * fp;
        dotgp += g[c] * v[c];
        if (c == t) wt = wcl[c];
      }
    const float cescale = gce * wt * inv_ce_den;
#pragma unroll
    for (int c = 0; c < SL_MAX_NC; ++c)
      if (c < nc) gz[(size_t)c * HW + p] = v[c] * (g[c] - dotgp) + cescale * (v[c] - (c == t ? 1.f : 0.f));
  }
}

// (bs, nc, HW) dense one-hot (f32 / i64 / u8) -> u8 class map (first maximum, like Tensor.max(1)[1])
template <typename T>
__global__ void onehot_argmax_kernel(const T* __restrict__ tgt, int nc, long long HW, long long total,
                                     unsigned char* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / HW, p = i % HW;
  const T* t = tgt + (size_t)b * nc * HW + p;
  T best = t[0];
  int arg = 0;
  for (int c = 1; c < nc; ++c) {
    const T v = t[(size_t)c * HW];
    if (v > best) { best = v; arg = c; }
  }
  out[i] = (unsigned char)arg;
}

}  // namespace

extern "C" size_t isa_seg_losses_workspace_bytes(int bs, int n_classes, long long HW) {
  if (bs <= 0 || n_classes <= 0 || HW <= 0) return 0;
  return seg_layout(bs, n_classes, HW).total;
}

extern "C" int isa_seg_losses_fwd(const float* logits, const unsigned char* class_map, const float* class_weights,
                                  int bs, int n_classes, long long HW, int dice_time, float smooth, int optimize_bg,
                                  float* out_ce_dice, void* workspace, size_t workspace_bytes, void* stream) {
  ISA_CHECK_ARG(logits && class_map && out_ce_dice && workspace, "isa_seg_losses_fwd: null pointer");
  ISA_CHECK_ARG(bs > 0 && HW > 0 && n_classes >= 2 && n_classes <= SL_MAX_NC,
                "isa_seg_losses_fwd: bs=%d HW=%lld n_classes=%d (2..%d supported)", bs, HW, n_classes, SL_MAX_NC);
  ISA_CHECK_ARG(dice_time == 1 || dice_time == 2, "isa_seg_losses_fwd: dice_time must be 1 or 2");
  ISA_CHECK_ARG(smooth > 0.f, "isa_seg_losses_fwd: smooth must be > 0 (dice.py:22)");
  const SegLayout L = seg_layout(bs, n_classes, HW);
  if (workspace_bytes < L.total) {
    isa_set_error("isa_seg_losses_fwd: workspace %zu < %zu", workspace_bytes, L.total);
    return ISA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ISA_CUDA(cudaMemsetAsync(workspace, 0, 16, st));
  dim3 grid(L.chunks, bs);
  seg_losses_fwd_kernel<<<grid, SL_THREADS, 0, st>>>(logits, class_map, class_weights, bs, n_classes, HW,
                                                     dice_time == 2, smooth, optimize_bg, (unsigned char*)workspace,
                                                     out_ce_dice);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

extern "C" int isa_seg_losses_bwd(const float* logits, const unsigned char* class_map, const float* class_weights,
                                  int bs, int n_classes, long long HW, int dice_time, const float* grad_ce,
                                  const float* grad_dice, float* grad_logits, const void* workspace,
                                  size_t workspace_bytes, void* stream) {
  ISA_CHECK_ARG(logits && class_map && grad_logits && workspace, "isa_seg_losses_bwd: null pointer");
  ISA_CHECK_ARG(bs > 0 && HW > 0 && n_classes >= 2 && n_classes <= SL_MAX_NC, "isa_seg_losses_bwd: bad shape");
  const SegLayout L = seg_layout(bs, n_classes, HW);
  if (workspace_bytes < L.total) {
    isa_set_error("isa_seg_losses_bwd: workspace %zu < %zu", workspace_bytes, L.total);
    return ISA_ERR_WORKSPACE;
  }
  dim3 grid(L.chunks, bs);
  seg_losses_bwd_kernel<<<grid, SL_THREADS, 0, (cudaStream_t)stream>>>(
      logits, class_map, class_weights, bs, n_classes, HW, dice_time == 2, (const unsigned char*)workspace, grad_ce,
      grad_dice, grad_logits);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

extern "C" int isa_onehot_argmax(const void* target, int target_kind, int bs, int n_classes, long long HW,
                                 unsigned char* class_map, void* stream) {
  ISA_CHECK_ARG(target && class_map, "isa_onehot_argmax: null pointer");
  ISA_CHECK_ARG(bs > 0 && HW > 0 && n_classes >= 1 && n_classes <= 255, "isa_onehot_argmax: bad shape");
  const long long total = (long long)bs * HW;
  const int th = 256;
  const unsigned blocks = (unsigned)((total + th - 1) / th);
  cudaStream_t st = (cudaStream_t)stream;
  switch (target_kind) {
    case 1: onehot_argmax_kernel<float><<<blocks, th, 0, st>>>((const float*)target, n_classes, HW, total, class_map); break;
    case 2: onehot_argmax_kernel<long long><<<blocks, th, 0, st>>>((const long long*)target, n_classes, HW, total, class_map); break;
    case 3: onehot_argmax_kernel<unsigned char><<<blocks, th, 0, st>>>((const unsigned char*)target, n_classes, HW, total, class_map); break;
    default:
      isa_set_error("isa_onehot_argmax: target_kind %d (1 f32, 2 i64, 3 u8)", target_kind);
      return ISA_ERR_BAD_ARG;
  }
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}
