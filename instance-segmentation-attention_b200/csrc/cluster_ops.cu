// The data movement either side of the clustering kernel, kept on the device:
//   isa_fg_compact              argmax over the semantic map, foreground compaction of the
//                               embedding (order of np.where, i.e. row-major)
//                               /root/reference/code/lib/prediction.py:57-69
//   isa_scatter_labels_upsample mask[y,x] = label+1 at foreground pixels, then
//                               cv2.resize(INTER_NEAREST) of both masks to the raw image size
//                               /root/reference/code/lib/prediction.py:76-83 and :47-50,105-108
// Integer / index work: results are bit-exact with numpy + OpenCV (tests/test_cluster_gpu.py).
#include "isa_common.cuh"

namespace {

constexpr int kBlk = 1024;  // pixels per compaction block

// class of a pixel = np.argmax over the class axis (first maximum wins)
__device__ __forceinline__ int argmax_class(const float* __restrict__ sem, int ncls, int HW, int p) {
  int best = 0;
  float bv = __ldg(sem + p);
  for (int c = 1; c < ncls; ++c) {
    const float v = __ldg(sem + (size_t)c * HW + p);
    if (v > bv) { bv = v; best = c; }
  }
  return best;
}

__global__ void __launch_bounds__(kBlk) fg_count_kernel(const float* __restrict__ sem, int ncls, int HW,
                                                        unsigned char* __restrict__ cls_map, int* __restrict__ block_counts) {
  const int p = blockIdx.x * kBlk + threadIdx.x;
  int cls = 0;
  if (p < HW) { cls = argmax_class(sem, ncls, HW, p); cls_map[p] = (unsigned char)cls; }
  const int c = __syncthreads_count(p < HW && cls != 0);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) fg_scan_kernel(int* __restrict__ block_counts, int nblocks, int* __restrict__ n_out) {
  // exclusive scan of block_counts in place (single CTA; nblocks <= a few thousand)
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nblocks) ? block_counts[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    if (lane == 31) s_warp[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int excl = s - v + (warp ? s_warp[warp - 1] : 0) + s_carry;
    if (i < nblocks) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = s_carry;
}

__global__ void __launch_bounds__(kBlk) fg_scatter_kernel(const float* __restrict__ emb, const unsigned char* __restrict__ cls_map,
                                                          const int* __restrict__ block_offsets, int C, int HW, int ld,
                                                          float* __restrict__ Xt, int* __restrict__ fg_index) {
  __shared__ int s_warp[32];
  const int p = blockIdx.x * kBlk + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fg = (p < HW) && cls_map[p] != 0;
  const unsigned bal = __ballot_sync(0xffffffffu, fg);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    const int v = s_warp[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    s_warp[lane] = s - v;
  }
  __syncthreads();
  if (fg) {
    const int o = block_offsets[blockIdx.x] + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
    fg_index[o] = p;
    for (int f = 0; f < C; ++f) Xt[(size_t)f * ld + o] = __ldg(emb + (size_t)f * HW + p);
  }
}

__global__ void zero_u8_kernel(unsigned char* __restrict__ m, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m[i] = 0;
}

__global__ void scatter_labels_kernel(const int* __restrict__ labels, const int* __restrict__ fg_index, const int* __restrict__ n_ptr,
                                      unsigned char* __restrict__ mask) {
  const int n = *n_ptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    mask[fg_index[i]] = (unsigned char)(labels[i] + 1);
}

// cv2.resize(INTER_NEAREST): sx = min(floor(dx * ifx), src_w - 1), ifx = 1/(dst_w/src_w) in double
__global__ void upsample_nearest2_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                         int src_h, int src_w, int dst_h, int dst_w, double ify, double ifx,
                                         unsigned char* __restrict__ oa, unsigned char* __restrict__ ob) {
  const int total = dst_h * dst_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dy = i / dst_w, dx = i % dst_w;
    const int sy = min(__double2int_rd(__dmul_rn((double)dy, ify)), src_h - 1);
    const int sx = min(__double2int_rd(__dmul_rn((double)dx, ifx)), src_w - 1);
    const int s = sy * src_w + sx;
    if (oa) oa[i] = a[s];
    if (ob) ob[i] = b[s];
  }
}

}  // namespace

extern "C" {

size_t isa_fg_compact_workspace_bytes(int HW) {
  if (HW <= 0) return 0;
  return isa_align_up(sizeof(int) * (size_t)((HW + kBlk - 1) / kBlk), 256);
}

// sem [ncls][HW] f32, emb [C][HW] f32 -> cls_map [HW] u8 (argmax class), Xt [C][ld] f32 (first n columns),
// fg_index [HW] i32 (first n entries: pixel index of each foreground point), n_out [1] (device).
int isa_fg_compact(const float* sem, const float* emb, int ncls, int C, int HW, int ld,
                   unsigned char* cls_map, float* Xt, int* fg_index, int* n_out,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(sem && emb && cls_map && Xt && fg_index && n_out && workspace, "fg_compact: null pointer");
  ISA_CHECK_ARG(ncls >= 1 && ncls <= 255 && C > 0 && HW > 0 && ld >= HW, "fg_compact: bad dimensions (ncls=%d C=%d HW=%d ld=%d)", ncls, C, HW, ld);
  if (workspace_bytes < isa_fg_compact_workspace_bytes(HW)) {
    isa_set_error("fg_compact: workspace too small");
    return ISA_ERR_WORKSPACE;
  }
  const int nb = (HW + kBlk - 1) / kBlk;
  int* counts = (int*)workspace;
  fg_count_kernel<<<nb, kBlk, 0, stream>>>(sem, ncls, HW, cls_map, counts);
  fg_scan_kernel<<<1, 1024, 0, stream>>>(counts, nb, n_out);
  fg_scatter_kernel<<<nb, kBlk, 0, stream>>>(emb, cls_map, counts, C, HW, ld, Xt, fg_index);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

// labels [n] i32 + fg_index [n] -> ins_small [h][w] u8 (0 = background, label+1 otherwise);
// then nearest up-sampling of ins_small and cls_map to (out_h, out_w).  Either output may be NULL.
int isa_scatter_labels_upsample(const int* labels, const int* fg_index, const int* n_ptr,
                                const unsigned char* cls_map, int h, int w, int out_h, int out_w,
                                unsigned char* ins_small, unsigned char* ins_up, unsigned char* cls_up,
                                cudaStream_t stream) {
  ISA_CHECK_ARG(labels && fg_index && n_ptr && ins_small, "scatter_labels_upsample: null pointer");
  ISA_CHECK_ARG(h > 0 && w > 0 && out_h > 0 && out_w > 0, "scatter_labels_upsample: bad sizes");
  const int HW = h * w;
  int grid = (HW + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  zero_u8_kernel<<<grid, 256, 0, stream>>>(ins_small, HW);
  scatter_labels_kernel<<<grid, 256, 0, stream>>>(labels, fg_index, n_ptr, ins_small);
  if (ins_up || cls_up) {
    ISA_CHECK_ARG(!cls_up || cls_map, "scatter_labels_upsample: cls_up requested without cls_map");
    const double ify = 1.0 / ((double)out_h / (double)h);
    const double ifx = 1.0 / ((double)out_w / (double)w);
    int g2 = (out_h * out_w + 255) / 256;
    if (g2 > 148 * 8) g2 = 148 * 8;
    upsample_nearest2_kernel<<<g2, 256, 0, stream>>>(ins_small, cls_map, h, w, out_h, out_w, ify, ifx, ins_up, cls_up);
  }
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}

}  // extern "C"
