// fp32 -> 3 x bf16 operand split for the ReNet projections.
//
// The input / weight-gradient projections of the ReNet sweeps (x W_ih^T over all tokens, dG W_ih,
// dG^T [x | h_prev | 1]) are plain GEMMs; on fp32 CUDA cores they cost more than the recurrent scan
// itself (r1b launch list: ~8.5 ms of the 31.6 ms training step in `simt_sgemm`).  They run on the
// bf16 tensor pipe at fp32-level accuracy instead: each operand is split into bf16 hi + lo parts and
//     a b  ~=  a_hi b_hi + a_hi b_lo + a_lo b_hi          (error ~2^-16 relative, fp32 accumulate)
// is evaluated as ONE bf16 GEMM whose K dimension concatenates the three pairings.  This kernel
// writes the three parts of a row-major fp32 matrix in one pass:
//     order 0 (left operand):  parts (hi, hi, lo)        order 1 (right operand): parts (hi, lo, hi)
// part q of element (r, c) goes to dst[q * part_stride + r * dst_ld + c], so the same kernel produces
// the K-concatenated layout ([rows][3][cols]: dst_ld = 3 cols, part_stride = cols) and column blocks
// of a wider matrix.  `shift` reads row r + shift instead of r and writes zeros where the shifted row
// leaves its sequence (position (r / pos_div) % pos_mod): that is how h_{t-1} of a sweep direction is
// taken straight from the forward output without materialising a shifted copy
// (the GRU contract: /root/reference/code/lib/archs/modules/README.md:225-256).
#include "isa_common.cuh"
#include <cuda_bf16.h>

namespace {

struct SplitParams {
  const float* src;
  __nv_bfloat16* dst;
  long long rows, src_ld, dst_ld, part_stride, shift;
  int cols, order, pos_div, pos_mod, pos_step;
};

__device__ __forceinline__ void split1(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// one thread = 4 consecutive columns of one row (cols % 4 == 0, 16 B aligned rows)
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const SplitParams p) {
  const int c4 = p.cols >> 2;
  const long long total = p.rows * c4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / c4;
    const int c = (int)(idx - r * c4) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = true;
    if (p.pos_step != 0) {
      const int pos = (int)((r / p.pos_div) % p.pos_mod) + p.pos_step;
      ok = pos >= 0 && pos < p.pos_mod;
    }
    if (ok) v = __ldg(reinterpret_cast<const float4*>(p.src + (r + p.shift) * p.src_ld + c));
    __nv_bfloat16 h[4], l[4];
    split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
    uint2 hv, lv;
    hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    __nv_bfloat16* d0 = p.dst + r * p.dst_ld + c;
    *reinterpret_cast<uint2*>(d0) = hv;
    *reinterpret_cast<uint2*>(d0 + p.part_stride) = p.order == 0 ? hv : lv;
    *reinterpret_cast<uint2*>(d0 + 2 * p.part_stride) = p.order == 0 ? lv : hv;
  }
}

// cols / 4 <= 256: thread t < A (A = largest multiple of cols/4 <= 256) owns column group t % (cols/4) for the whole
// kernel and walks the rows with a fixed stride -- no per-element index division (the generic kernel below pays a 64-bit
// division per 16 bytes), two rows in flight per thread
__global__ void __launch_bounds__(256) split_bf16x3_rows_kernel(const SplitParams p) {
  const int c4 = p.cols >> 2;
  const int rpi = 256 / c4;
  const int t = threadIdx.x;
  if (t >= rpi * c4) return;
  const int c = (t % c4) << 2;
  const long long rstep = (long long)gridDim.x * rpi;
  auto load_row = [&](long long r) -> float4 {
    bool ok = true;
    if (p.pos_step != 0) {
      const int pos = (int)((r / p.pos_div) % p.pos_mod) + p.pos_step;
      ok = pos >= 0 && pos < p.pos_mod;
    }
    return ok ? __ldg(reinterpret_cast<const float4*>(p.src + (r + p.shift) * p.src_ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_row = [&](long long r, const float4 v) {
    __nv_bfloat16 h[4], l[4];
    split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
    uint2 hv, lv;
    hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    __nv_bfloat16* d0 = p.dst + r * p.dst_ld + c;
    *reinterpret_cast<uint2*>(d0) = hv;
    *reinterpret_cast<uint2*>(d0 + p.part_stride) = p.order == 0 ? hv : lv;
    *reinterpret_cast<uint2*>(d0 + 2 * p.part_stride) = p.order == 0 ? lv : hv;
  };
  long long r = (long long)blockIdx.x * rpi + t / c4;
  // four independent rows per iteration: the loads are issued together before any store
  for (; r + 3 * rstep < p.rows; r += 4 * rstep) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = load_row(r + j * rstep);
#pragma unroll
    for (int j = 0; j < 4; ++j) store_row(r + j * rstep, v[j]);
  }
  for (; r < p.rows; r += rstep) store_row(r, load_row(r));
}

}  // namespace

extern "C" int isa_split_bf16x3(const float* src, long long rows, int cols, long long src_ld, void* dst, long long dst_ld,
                                long long part_stride, int order, long long shift, int pos_div, int pos_mod, int pos_step,
                                cudaStream_t stream) {
  ISA_CHECK_ARG(src && dst, "split_bf16x3: null pointer");
  ISA_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0, "split_bf16x3: cols must be a positive multiple of 4 (got %d)", cols);
  ISA_CHECK_ARG(src_ld % 4 == 0 && dst_ld % 4 == 0 && part_stride % 4 == 0, "split_bf16x3: leading dimensions must be multiples of 4");
  ISA_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0,
                "split_bf16x3: src must be 16 B and dst 8 B aligned");
  ISA_CHECK_ARG(order == 0 || order == 1, "split_bf16x3: order must be 0 (hi,hi,lo) or 1 (hi,lo,hi)");
  ISA_CHECK_ARG(pos_step == 0 || (pos_div > 0 && pos_mod > 0), "split_bf16x3: pos_div / pos_mod must be positive with a shift");
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  SplitParams p;
  p.src = src; p.dst = reinterpret_cast<__nv_bfloat16*>(dst);
  p.rows = rows; p.src_ld = src_ld; p.dst_ld = dst_ld; p.part_stride = part_stride; p.shift = shift;
  p.cols = cols; p.order = order; p.pos_div = pos_div; p.pos_mod = pos_mod; p.pos_step = pos_step;
  const long long cap = (long long)di.num_sms * 16;
  if (cols / 4 <= 256) {
    const int rpi = 256 / (cols / 4);
    long long blocks = (rows + rpi - 1) / rpi;
    if (blocks > cap) blocks = cap;
    split_bf16x3_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
  } else {
    const long long total = rows * (cols / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
  }
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}
