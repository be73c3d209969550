// Persistent bidirectional-GRU scan for the ReNet row / column sweeps.
//
// Reference contract: /root/reference/code/lib/archs/modules/README.md:225-256 (ReNet: a
// bidirectional GRU over every row, then over every column of the result; renet.py itself is
// absent from the tree, so the arithmetic is PyTorch's nn.GRU:
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
// with weight_hh_l0 laid out (3n, n) in gate order r, z, n).
//
// Split of the work: the input projection gx = x W_ih^T + b_ih is one big GEMM over all
// tokens (library GEMM on the host side); this kernel is the sequential part.  One CTA owns S
// sequences of ONE direction for the whole sweep:
//   * W_hh^T (n x 3n fp32, 120 KB for n = 100) is loaded into shared memory once and stays
//     there for all T steps; h_{t-1} of the S sequences lives in shared memory (double buffered);
//   * per step each thread owns one hidden unit j and S/2 sequences: 3 * S/2 fp32 accumulators
//     over i = 0..n-1 (W row reads are conflict free, h reads are warp broadcasts), then the
//     gate math in registers; the next step's gx values are prefetched into registers while the
//     current step computes;
//   * all tensors are token-major ("channels last"): a sequence's per-step chunk is contiguous,
//     so every global access of a warp is a full 128 B line.  Rows and columns differ only in
//     the token strides handed in.
// Backward (BPTT) walks the other way with W_hh (n-major rows) in shared memory:
//   dh_{t-1} = z * dh + W_hh^T [dr_pre, dz_pre, r * dn_pre]; it writes the gradient of the input
//   projection (dgx) and the hidden-side n-gate gradient; the weight gradients are GEMMs over
//   those arrays on the host side.
#include "isa_common.cuh"
#include <math.h>

namespace {

struct TokMap {
  int inner;
  long long outer_stride, inner_stride, t_stride;  // in tokens
};
__device__ __forceinline__ long long tok_base(const TokMap& m, int q) {
  return (long long)(q / m.inner) * m.outer_stride + (long long)(q % m.inner) * m.inner_stride;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

struct GruFwdParams {
  const float* gx;     // [tokens][2][3n]
  const float* w_hh;   // [2][3n][n]
  const float* b_hh;   // [2][3n]
  float* out;          // [tokens][2n]
  float* stash;        // [tokens][2][4n] (r, z, n, hn) or null
  int n_seq, T, n;
  TokMap map;
};

// S sequences per CTA, ST = S/2 per thread; thread = (unit j, sequence half sg)
template <int S>
__global__ void __launch_bounds__(256, 1) gru_scan_fwd_kernel(const GruFwdParams prm) {
  constexpr int ST = S / 2;
  extern __shared__ __align__(16) float smem[];
  const int n = prm.n, n3 = 3 * n;
  float* s_wt = smem;                   // [n][3n]   W_hh^T
  float* s_h = s_wt + (size_t)n * n3;   // [2][n][S]
  const int d = blockIdx.y;
  const int q0 = blockIdx.x * S;
  const int T = prm.T;

  // W_hh^T -> smem (one-time; coalesced global reads along i)
  const float* __restrict__ w = prm.w_hh + (size_t)d * n3 * n;
  for (int idx = threadIdx.x; idx < n3 * n; idx += blockDim.x) {
    const int g = idx / n, i = idx % n;
    s_wt[(size_t)i * n3 + g] = __ldg(w + idx);
  }
  for (int idx = threadIdx.x; idx < 2 * n * S; idx += blockDim.x) s_h[idx] = 0.f;

  const int j = threadIdx.x % n;     // hidden unit (threads >= 2n idle in the math)
  const int sg = threadIdx.x / n;    // 0 or 1
  const bool active = threadIdx.x < 2 * n;
  const float bhr = active ? __ldg(prm.b_hh + d * n3 + j) : 0.f;
  const float bhz = active ? __ldg(prm.b_hh + d * n3 + n + j) : 0.f;
  const float bhn = active ? __ldg(prm.b_hh + d * n3 + 2 * n + j) : 0.f;

  long long base[ST];
  bool valid[ST];
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    const int q = q0 + sg * ST + s;
    valid[s] = active && q < prm.n_seq;
    base[s] = valid[s] ? tok_base(prm.map, q) : 0;
  }
  const float* __restrict__ gx = prm.gx + (size_t)d * n3 + j;

  float gr[ST], gz[ST], gn[ST];  // prefetched x-projection of the coming step
  auto prefetch = [&](int t) {
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      if (valid[s]) {
        const float* p = gx + (size_t)(base[s] + (long long)t * prm.map.t_stride) * (2 * n3);
        gr[s] = __ldg(p); gz[s] = __ldg(p + n); gn[s] = __ldg(p + 2 * n);
      } else { gr[s] = gz[s] = gn[s] = 0.f; }
    }
  };
  prefetch(d == 0 ? 0 : T - 1);
  __syncthreads();

  for (int step = 0; step < T; ++step) {
    const int t = (d == 0) ? step : T - 1 - step;
    const float* __restrict__ hc = s_h + (size_t)(step & 1) * n * S;
    float* __restrict__ hn_buf = s_h + (size_t)((step + 1) & 1) * n * S;
    float xr[ST], xz[ST], xn[ST];
#pragma unroll
    for (int s = 0; s < ST; ++s) { xr[s] = gr[s]; xz[s] = gz[s]; xn[s] = gn[s]; }
    if (step + 1 < T) prefetch(d == 0 ? step + 1 : T - 2 - step);

    if (active) {
      float ar[ST], az[ST], an[ST];
#pragma unroll
      for (int s = 0; s < ST; ++s) { ar[s] = 0.f; az[s] = 0.f; an[s] = 0.f; }
      const float* __restrict__ wp = s_wt + j;
      const float* __restrict__ hp = hc + sg * ST;
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const float wr = wp[(size_t)i * n3], wz = wp[(size_t)i * n3 + n], wn = wp[(size_t)i * n3 + 2 * n];
        float hv[ST];
#pragma unroll
        for (int s4 = 0; s4 < ST; s4 += 4) {
          if (ST >= 4) {
            const float4 v = *reinterpret_cast<const float4*>(hp + (size_t)i * S + s4);
            hv[s4] = v.x; hv[s4 + 1] = v.y; hv[s4 + 2] = v.z; hv[s4 + 3] = v.w;
          }
        }
        if (ST < 4) {
#pragma unroll
          for (int s = 0; s < ST; ++s) hv[s] = hp[(size_t)i * S + s];
        }
#pragma unroll
        for (int s = 0; s < ST; ++s) {
          ar[s] = fmaf(wr, hv[s], ar[s]);
          az[s] = fmaf(wz, hv[s], az[s]);
          an[s] = fmaf(wn, hv[s], an[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        const float hprev = hc[(size_t)j * S + sg * ST + s];
        const float r = sigmoidf_acc(xr[s] + ar[s] + bhr);
        const float z = sigmoidf_acc(xz[s] + az[s] + bhz);
        const float hnn = an[s] + bhn;
        const float nn = tanhf(xn[s] + r * hnn);
        const float hnew = (1.f - z) * nn + z * hprev;
        hn_buf[(size_t)j * S + sg * ST + s] = hnew;
        if (valid[s]) {
          const long long tok = base[s] + (long long)t * prm.map.t_stride;
          prm.out[(size_t)tok * (2 * n) + d * n + j] = hnew;
          if (prm.stash) {
            float* st = prm.stash + ((size_t)tok * 2 + d) * (4 * n) + j;
            st[0] = r; st[n] = z; st[2 * n] = nn; st[3 * n] = hnn;
          }
        }
      }
    }
    __syncthreads();
  }
}

struct GruBwdParams {
  const float* dout;   // [tokens][2n]
  const float* out;    // [tokens][2n]  forward output (h_t)
  const float* stash;  // [tokens][2][4n]
  const float* w_hh;   // [2][3n][n]
  float* dgx;          // [tokens][2][3n]
  float* dghn;         // [tokens][2][n]   r * dn_pre (hidden-side n-gate gradient)
  int n_seq, T, n;
  TokMap map;
};

template <int S>
__global__ void __launch_bounds__(256, 1) gru_scan_bwd_kernel(const GruBwdParams prm) {
  constexpr int ST = S / 2;
  extern __shared__ __align__(16) float smem[];
  const int n = prm.n, n3 = 3 * n;
  float* s_w = smem;                     // [3n][n]  W_hh as stored
  float* s_dg = s_w + (size_t)n3 * n;    // [3n][S]  hidden-side gate gradients of this step
  float* s_dh = s_dg + (size_t)n3 * S;   // [n][S]   dh carried to the previous step
  const int d = blockIdx.y;
  const int q0 = blockIdx.x * S;
  const int T = prm.T;
  const float* __restrict__ w = prm.w_hh + (size_t)d * n3 * n;
  for (int idx = threadIdx.x; idx < n3 * n; idx += blockDim.x) s_w[idx] = __ldg(w + idx);
  for (int idx = threadIdx.x; idx < n * S; idx += blockDim.x) s_dh[idx] = 0.f;

  const int j = threadIdx.x % n;
  const int sg = threadIdx.x / n;
  const bool active = threadIdx.x < 2 * n;
  long long base[ST];
  bool valid[ST];
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    const int q = q0 + sg * ST + s;
    valid[s] = active && q < prm.n_seq;
    base[s] = valid[s] ? tok_base(prm.map, q) : 0;
  }
  __syncthreads();

  // backward visits the steps in the opposite order of the forward sweep of this direction
  for (int step = T - 1; step >= 0; --step) {
    const int t = (d == 0) ? step : T - 1 - step;        // token position of forward step `step`
    const int tp = (d == 0) ? t - 1 : t + 1;             // position holding h_{step-1}
    if (active) {
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        float dr_pre = 0.f, dz_pre = 0.f, dn_pre = 0.f, dnr = 0.f, dh_direct = 0.f;
        if (valid[s]) {
          const long long tok = base[s] + (long long)t * prm.map.t_stride;
          const float* st = prm.stash + ((size_t)tok * 2 + d) * (4 * n) + j;
          const float r = __ldg(st), z = __ldg(st + n), nn = __ldg(st + 2 * n), hnn = __ldg(st + 3 * n);
          float hprev = 0.f;
          if (step > 0) hprev = __ldg(prm.out + (size_t)(base[s] + (long long)tp * prm.map.t_stride) * (2 * n) + d * n + j);
          const float dh = __ldg(prm.dout + (size_t)tok * (2 * n) + d * n + j) + s_dh[(size_t)j * S + sg * ST + s];
          const float dn = dh * (1.f - z);
          const float dz = dh * (hprev - nn);
          dh_direct = dh * z;
          dn_pre = dn * (1.f - nn * nn);
          dz_pre = dz * z * (1.f - z);
          dr_pre = dn_pre * hnn * r * (1.f - r);
          dnr = dn_pre * r;
          float* g = prm.dgx + ((size_t)tok * 2 + d) * n3 + j;
          g[0] = dr_pre; g[n] = dz_pre; g[2 * n] = dn_pre;
          prm.dghn[((size_t)tok * 2 + d) * n + j] = dnr;
        }
        s_dg[(size_t)j * S + sg * ST + s] = dr_pre;
        s_dg[(size_t)(n + j) * S + sg * ST + s] = dz_pre;
        s_dg[(size_t)(2 * n + j) * S + sg * ST + s] = dnr;
        // stash the direct term in registers via smem slot reuse below
        s_dh[(size_t)j * S + sg * ST + s] = dh_direct;
      }
    }
    __syncthreads();
    if (active) {
      // dh_{step-1}[i=j] = z*dh + sum_g W_hh[g][j] * dg[g]
      float acc[ST];
#pragma unroll
      for (int s = 0; s < ST; ++s) acc[s] = 0.f;
      const float* __restrict__ wp = s_w + j;
      const float* __restrict__ gp = s_dg + sg * ST;
#pragma unroll 4
      for (int g = 0; g < n3; ++g) {
        const float wv = wp[(size_t)g * n];
        float gv[ST];
#pragma unroll
        for (int s4 = 0; s4 < ST; s4 += 4) {
          if (ST >= 4) {
            const float4 v = *reinterpret_cast<const float4*>(gp + (size_t)g * S + s4);
            gv[s4] = v.x; gv[s4 + 1] = v.y; gv[s4 + 2] = v.z; gv[s4 + 3] = v.w;
          }
        }
        if (ST < 4) {
#pragma unroll
          for (int s = 0; s < ST; ++s) gv[s] = gp[(size_t)g * S + s];
        }
#pragma unroll
        for (int s = 0; s < ST; ++s) acc[s] = fmaf(wv, gv[s], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < ST; ++s) s_dh[(size_t)j * S + sg * ST + s] += acc[s];
    }
    __syncthreads();
  }
}

size_t fwd_smem(int n, int S) { return sizeof(float) * ((size_t)n * 3 * n + 2 * (size_t)n * S); }
size_t bwd_smem(int n, int S) { return sizeof(float) * ((size_t)n * 3 * n + 3 * (size_t)n * S + (size_t)n * S); }

int pick_S(int n_seq, int num_sms) {
  // fill the machine: prefer the largest S that still gives >= ~0.8 waves of CTAs (2 directions)
  if ((long long)((n_seq + 15) / 16) * 2 * 10 >= (long long)num_sms * 8) return 16;
  if ((long long)((n_seq + 7) / 8) * 2 * 10 >= (long long)num_sms * 8) return 8;
  return 4;
}

int check_gru(int n_seq, int T, int n, int inner) {
  ISA_CHECK_ARG(n_seq > 0 && T > 0, "gru_scan: n_seq and T must be positive");
  ISA_CHECK_ARG(n >= 4 && n <= 128 && n % 4 == 0, "gru_scan: n_units must be a multiple of 4 in [4,128] (got %d)", n);
  ISA_CHECK_ARG(inner > 0, "gru_scan: inner must be positive");
  ISA_CHECK_ARG((n_seq + 3) / 4 <= 65535 * 32, "gru_scan: too many sequences");
  return ISA_OK;
}

}  // namespace

extern "C" {

int isa_gru_scan_fwd(const float* gx, const float* w_hh, const float* b_hh, int n_seq, int T, int n_units,
                     int inner, long long outer_tok_stride, long long inner_tok_stride, long long t_tok_stride,
                     float* out, float* stash, cudaStream_t stream) {
  int rc = check_gru(n_seq, T, n_units, inner);
  if (rc) return rc;
  ISA_CHECK_ARG(gx && w_hh && b_hh && out, "gru_scan_fwd: null pointer");
  IsaDeviceInfo di;
  rc = isa_device_info(&di);
  if (rc) return rc;
  GruFwdParams prm;
  prm.gx = gx; prm.w_hh = w_hh; prm.b_hh = b_hh; prm.out = out; prm.stash = stash;
  prm.n_seq = n_seq; prm.T = T; prm.n = n_units;
  prm.map.inner = inner; prm.map.outer_stride = outer_tok_stride; prm.map.inner_stride = inner_tok_stride; prm.map.t_stride = t_tok_stride;
  const int S = pick_S(n_seq, di.num_sms);
  const size_t smem = fwd_smem(n_units, S);
  ISA_CHECK_ARG(smem <= (size_t)di.max_smem_optin, "gru_scan_fwd: n_units=%d needs %zu B of shared memory (> %d)", n_units, smem, di.max_smem_optin);
  const int threads = ((2 * n_units + 31) / 32) * 32;
  dim3 grid((n_seq + S - 1) / S, 2);
#define LAUNCH_FWD(SV)                                                                                          \
  {                                                                                                             \
    ISA_CUDA(cudaFuncSetAttribute(gru_scan_fwd_kernel<SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gru_scan_fwd_kernel<SV><<<grid, threads, smem, stream>>>(prm);                                              \
  }
  if (S == 16) LAUNCH_FWD(16) else if (S == 8) LAUNCH_FWD(8) else LAUNCH_FWD(4)
#undef LAUNCH_FWD
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

int isa_gru_scan_bwd(const float* dout, const float* out, const float* stash, const float* w_hh,
                     int n_seq, int T, int n_units,
                     int inner, long long outer_tok_stride, long long inner_tok_stride, long long t_tok_stride,
                     float* dgx, float* dghn, cudaStream_t stream) {
  int rc = check_gru(n_seq, T, n_units, inner);
  if (rc) return rc;
  ISA_CHECK_ARG(dout && out && stash && w_hh && dgx && dghn, "gru_scan_bwd: null pointer");
  IsaDeviceInfo di;
  rc = isa_device_info(&di);
  if (rc) return rc;
  GruBwdParams prm;
  prm.dout = dout; prm.out = out; prm.stash = stash; prm.w_hh = w_hh; prm.dgx = dgx; prm.dghn = dghn;
  prm.n_seq = n_seq; prm.T = T; prm.n = n_units;
  prm.map.inner = inner; prm.map.outer_stride = outer_tok_stride; prm.map.inner_stride = inner_tok_stride; prm.map.t_stride = t_tok_stride;
  const int S = pick_S(n_seq, di.num_sms);
  const size_t smem = bwd_smem(n_units, S);
  ISA_CHECK_ARG(smem <= (size_t)di.max_smem_optin, "gru_scan_bwd: n_units=%d needs %zu B of shared memory (> %d)", n_units, smem, di.max_smem_optin);
  const int threads = ((2 * n_units + 31) / 32) * 32;
  dim3 grid((n_seq + S - 1) / S, 2);
#define LAUNCH_BWD(SV)                                                                                          \
  {                                                                                                             \
    ISA_CUDA(cudaFuncSetAttribute(gru_scan_bwd_kernel<SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gru_scan_bwd_kernel<SV><<<grid, threads, smem, stream>>>(prm);                                              \
  }
  if (S == 16) LAUNCH_BWD(16) else if (S == 8) LAUNCH_BWD(8) else LAUNCH_BWD(4)
#undef LAUNCH_BWD
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
