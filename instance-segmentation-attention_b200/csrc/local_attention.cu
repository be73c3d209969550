// Local 3x3 dilated attention (the reference's long-context alternative to the dense L x L block):
//   /root/reference/code/lib/archs/modules/utils.py:267-303  _ScalePDAttention.forward
// per pixel p and head:  s_i = (K[p + o_i] . Q[p]) * c^-1/2,  i = 0..8,  o_i = ((i/3 - 1) d, (i%3 - 1) d)
//                        s_i = -inf where nomask[p + o_i] != 0;  P = softmax_9(s), NaN -> 0;  out[p] = sum_i P_i V[p + o_i]
// Out-of-image neighbours are the reference's zero padding: K = V = 0 and nomask = 0 there, i.e. they
// take part in the softmax with score 0 and contribute nothing to the output (utils.py:281-285).
// The reference builds nine shifted copies of K, V and the mask with torch.cat and runs b*h*w micro-bmm's;
// here it is a stencil: one thread per pixel, channel-planar (NCHW) operands so that a warp reads 128
// contiguous bytes per channel and neighbour, neighbours served by L1/L2.
// Backward is two gather passes (no atomics): (a) per pixel dP, dS -> dQ and the saved dS;
// (b) per pixel q: dK[q] = scale * sum_i dS_i[q - o_i] Q[q - o_i],  dV[q] = sum_i P_i[q - o_i] dO[q - o_i].
#include "isa_common.cuh"
#include <math.h>

namespace {

constexpr int kMaxD = 32;

struct LaParams {
  const float* Q; const float* K; const float* V; const float* nomask;  // [Bh][dk|dv|1][h][w]
  float* out; float* P;                                                  // [Bh][dv][h][w], [Bh][9][h][w]
  const float* dout; float* dQ; float* dS; float* dK; float* dV;
  int Bh, dk, dv, h, w, dil, mask_batches;
  float scale;
};

__device__ __forceinline__ bool nb(const LaParams& p, int y, int x, int i, int& yy, int& xx) {
  yy = y + (i / 3 - 1) * p.dil;
  xx = x + (i % 3 - 1) * p.dil;
  return yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
}

template <int DK>
__global__ void __launch_bounds__(128) la_fwd_kernel(const LaParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= hw) return;
  const int y = pix / p.w, x = pix % p.w;
  const float* Qb = p.Q + (size_t)b * p.dk * hw;
  const float* Kb = p.K + (size_t)b * p.dk * hw;
  const float* Vb = p.V + (size_t)b * p.dv * hw;
  const float* Mb = p.nomask ? p.nomask + (size_t)(b % p.mask_batches) * hw : nullptr;
  float q[DK];
#pragma unroll
  for (int c = 0; c < DK; ++c) q[c] = c < p.dk ? __ldg(Qb + (size_t)c * hw + pix) : 0.f;
  float s[9];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    int yy, xx;
    float v = 0.f;
    if (nb(p, y, x, i, yy, xx)) {
      const int np = yy * p.w + xx;
      if (Mb && __ldg(Mb + np) != 0.f) v = -INFINITY;
      else {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DK; ++c)
          if (c < p.dk) acc = fmaf(__ldg(Kb + (size_t)c * hw + np), q[c], acc);
        v = acc * p.scale;
      }
    }
    s[i] = v;
    m = fmaxf(m, v);
  }
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) { s[i] = (m == -INFINITY) ? 0.f : expf(s[i] - m); l += s[i]; }
  const float inv = l > 0.f ? 1.f / l : 0.f;   // all nine masked: softmax is NaN, replaced by 0 (utils.py:297)
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    s[i] *= inv;
    if (p.P) p.P[((size_t)b * 9 + i) * hw + pix] = s[i];
  }
  for (int c = 0; c < p.dv; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      int yy, xx;
      if (s[i] != 0.f && nb(p, y, x, i, yy, xx)) acc = fmaf(s[i], __ldg(Vb + (size_t)c * hw + yy * p.w + xx), acc);
    }
    p.out[((size_t)b * p.dv + c) * hw + pix] = acc;
  }
}

template <int DK>
__global__ void __launch_bounds__(128) la_bwd_a_kernel(const LaParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= hw) return;
  const int y = pix / p.w, x = pix % p.w;
  const float* Kb = p.K + (size_t)b * p.dk * hw;
  const float* Vb = p.V + (size_t)b * p.dv * hw;
  const float* Gb = p.dout + (size_t)b * p.dv * hw;
  float P[9], dP[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { P[i] = __ldg(p.P + ((size_t)b * 9 + i) * hw + pix); dP[i] = 0.f; }
  for (int c = 0; c < p.dv; ++c) {
    const float g = __ldg(Gb + (size_t)c * hw + pix);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      int yy, xx;
      if (nb(p, y, x, i, yy, xx)) dP[i] = fmaf(g, __ldg(Vb + (size_t)c * hw + yy * p.w + xx), dP[i]);
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) dot = fmaf(P[i], dP[i], dot);
  float dq[DK];
#pragma unroll
  for (int c = 0; c < DK; ++c) dq[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float ds = P[i] * (dP[i] - dot) * p.scale;   // gradient of the pre-scale score K.Q
    p.dS[((size_t)b * 9 + i) * hw + pix] = ds;
    int yy, xx;
    if (ds != 0.f && nb(p, y, x, i, yy, xx)) {
      const int np = yy * p.w + xx;
#pragma unroll
      for (int c = 0; c < DK; ++c)
        if (c < p.dk) dq[c] = fmaf(ds, __ldg(Kb + (size_t)c * hw + np), dq[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < DK; ++c)
    if (c < p.dk) p.dQ[((size_t)b * p.dk + c) * hw + pix] = dq[c];
}

// gather form: pixel q receives from every centre pc = q - o_i that has q as its neighbour i
__global__ void __launch_bounds__(128) la_bwd_b_kernel(const LaParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= hw) return;
  const int y = pix / p.w, x = pix % p.w;
  int cpix[9];
  float ds[9], pr[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int yy = y - (i / 3 - 1) * p.dil, xx = x - (i % 3 - 1) * p.dil;
    const bool ok = yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
    cpix[i] = ok ? yy * p.w + xx : -1;
    ds[i] = ok ? __ldg(p.dS + ((size_t)b * 9 + i) * hw + cpix[i]) : 0.f;
    pr[i] = ok ? __ldg(p.P + ((size_t)b * 9 + i) * hw + cpix[i]) : 0.f;
  }
  const float* Qb = p.Q + (size_t)b * p.dk * hw;
  const float* Gb = p.dout + (size_t)b * p.dv * hw;
  for (int c = 0; c < p.dk; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (ds[i] != 0.f) acc = fmaf(ds[i], __ldg(Qb + (size_t)c * hw + cpix[i]), acc);
    p.dK[((size_t)b * p.dk + c) * hw + pix] = acc;
  }
  for (int c = 0; c < p.dv; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (pr[i] != 0.f) acc = fmaf(pr[i], __ldg(Gb + (size_t)c * hw + cpix[i]), acc);
    p.dV[((size_t)b * p.dv + c) * hw + pix] = acc;
  }
}

int check_la(int Bh, int dk, int dv, int h, int w, int dil, int mask_batches) {
  ISA_CHECK_ARG(Bh > 0 && Bh <= 65535 && h > 0 && w > 0 && dil > 0, "local_attention: bad sizes");
  ISA_CHECK_ARG(dk > 0 && dk <= kMaxD && dv > 0, "local_attention: d_k must be in [1,%d] (got %d)", kMaxD, dk);
  ISA_CHECK_ARG(mask_batches > 0, "local_attention: mask_batches must be positive");
  return ISA_OK;
}

}  // namespace

extern "C" {

int isa_local_attention_fwd(const float* Q, const float* K, const float* V, const float* nomask, int mask_batches, int Bh, int dk, int dv,
                            int h, int w, int dil, float scale, float* out, float* P, cudaStream_t stream) {
  int rc = check_la(Bh, dk, dv, h, w, dil, mask_batches);
  if (rc) return rc;
  ISA_CHECK_ARG(Q && K && V && out, "local_attention_fwd: null pointer");
  LaParams p = {};
  p.Q = Q; p.K = K; p.V = V; p.nomask = nomask; p.out = out; p.P = P;
  p.Bh = Bh; p.dk = dk; p.dv = dv; p.h = h; p.w = w; p.dil = dil; p.mask_batches = mask_batches; p.scale = scale;
  dim3 grid((h * w + 127) / 128, Bh);
  if (dk <= 16) la_fwd_kernel<16><<<grid, 128, 0, stream>>>(p);
  else la_fwd_kernel<kMaxD><<<grid, 128, 0, stream>>>(p);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_local_attention_bwd(const float* Q, const float* K, const float* V, const float* P, const float* dout, int Bh, int dk, int dv,
                            int h, int w, int dil, float scale, float* dQ, float* dK, float* dV, float* dS_workspace, cudaStream_t stream) {
  int rc = check_la(Bh, dk, dv, h, w, dil, 1);
  if (rc) return rc;
  ISA_CHECK_ARG(Q && K && V && P && dout && dQ && dK && dV && dS_workspace, "local_attention_bwd: null pointer");
  LaParams p = {};
  p.Q = Q; p.K = K; p.V = V; p.P = const_cast<float*>(P); p.dout = dout; p.dQ = dQ; p.dK = dK; p.dV = dV; p.dS = dS_workspace;
  p.Bh = Bh; p.dk = dk; p.dv = dv; p.h = h; p.w = w; p.dil = dil; p.mask_batches = 1; p.scale = scale;
  dim3 grid((h * w + 127) / 128, Bh);
  if (dk <= 16) la_bwd_a_kernel<16><<<grid, 128, 0, stream>>>(p);
  else la_bwd_a_kernel<kMaxD><<<grid, 128, 0, stream>>>(p);
  ISA_CUDA(cudaGetLastError());
  la_bwd_b_kernel<<<grid, 128, 0, stream>>>(p);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
