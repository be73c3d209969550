// Symmetric best dice (SBD) and |DiC| of instance label images on the device.  Replaces the O(n_gt * n_pred * H*W)
// numpy loops of /root/reference/code/evaluate.py:18-57 (calc_dice, calc_bd, calc_sbd, calc_dic) with ONE pass over
// the two uint8 label images (a 256 x 256 contingency table per image, warp-aggregated integer atomics: exact and
// order independent) and a per-image reduction of the table:
//   dice(i, j) = 2 |gt == i and pred == j| / (|gt == i| + |pred == j|)
//   BD(a, b)   = mean over the objects i != 0 of a of max over the objects j != 0 of b of dice(i, j)
//   SBD        = min(BD(gt, pred), BD(pred, gt));   |DiC| = |#objects(gt) - #objects(pred)|
// HBM-bound: 2 bytes read per pixel.  Every dice is the same ratio of integer counts as in the reference, evaluated in
// float64, so the values are identical.
#include "isa_common.cuh"
#include <math.h>

namespace {

__global__ void __launch_bounds__(256) sbd_table_kernel(const unsigned char* __restrict__ gt, const unsigned char* __restrict__ pred,
                                                        long long P, unsigned* __restrict__ table) {
  const int img = blockIdx.y;
  const unsigned char* a = gt + (size_t)img * P;
  const unsigned char* b = pred + (size_t)img * P;
  unsigned* t = table + (size_t)img * 65536;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    const unsigned key = ((unsigned)a[i] << 8) | (unsigned)b[i];
    // label images are spatially coherent: most lanes of a warp hold the same (gt, pred) pair -> one atomic per group
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, key);
    if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(t + key, (unsigned)__popc(grp));
  }
}

// numpy's float64 sum (np.mean of evaluate.py:45): sequential below 8 elements, else eight interleaved partial sums folded
// pairwise plus a sequential tail, halves recursively above 128 elements -- restated so the mean is bit-identical
__device__ double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += a[i];
    return r;
  }
  if (n <= 128) {
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// one CTA per image: out[img] = {SBD, BD(gt, pred), BD(pred, gt), n_objects(gt), n_objects(pred)}
__global__ void __launch_bounds__(256) sbd_reduce_kernel(const unsigned* __restrict__ table, double* __restrict__ out) {
  const int img = blockIdx.x, j = threadIdx.x;
  const unsigned* t = table + (size_t)img * 65536;
  __shared__ unsigned s_row[256], s_col[256];
  __shared__ double s_best[256];
  __shared__ double s_res[2];
  __shared__ int s_cnt[2];
  unsigned r = 0, c = 0;
  for (int q = 0; q < 256; ++q) { r += t[j * 256 + q]; c += t[q * 256 + j]; }
  s_row[j] = r;      // |gt == j|
  s_col[j] = c;      // |pred == j|
  __syncthreads();
  for (int dir = 0; dir < 2; ++dir) {
    // thread j = object j of the first image of this direction
    const unsigned mine = dir == 0 ? s_row[j] : s_col[j];
    double best = -1.0;
    if (j != 0 && mine > 0) {
      for (int q = 1; q < 256; ++q) {
        const unsigned other = dir == 0 ? s_col[q] : s_row[q];
        if (other == 0) continue;
        const unsigned inter = dir == 0 ? t[j * 256 + q] : t[q * 256 + j];
        const double d = 2.0 * (double)inter / ((double)mine + (double)other);
        if (d > best) best = d;
      }
    }
    s_best[j] = best;
    __syncthreads();
    if (j == 0) {
      int n = 0, empty_other = 0;
      for (int q = 1; q < 256; ++q) {        // compact in ascending label order (np.unique), in place: n <= q
        const unsigned m = dir == 0 ? s_row[q] : s_col[q];
        if (m == 0) continue;
        if (s_best[q] < 0.0) empty_other = 1;
        s_best[n++] = s_best[q];
      }
      // no object in the first image: np.mean([]) = nan; none in the second: np.max([]) raises in the reference -> nan here
      s_res[dir] = (n == 0 || empty_other) ? nan("") : np_pairwise_sum(s_best, n) / (double)n;
      s_cnt[dir] = n;
    }
    __syncthreads();
  }
  if (j == 0) {
    double* o = out + (size_t)img * 5;
    const double a = s_res[0], b = s_res[1];
    o[0] = (isnan(a) || isnan(b)) ? nan("") : (a < b ? a : b);
    o[1] = a; o[2] = b; o[3] = (double)s_cnt[0]; o[4] = (double)s_cnt[1];
  }
}

}  // namespace

extern "C" size_t isa_sbd_workspace_bytes(int n_images) { return n_images > 0 ? (size_t)n_images * 65536 * sizeof(unsigned) : 0; }

extern "C" int isa_sbd(const unsigned char* gt, const unsigned char* pred, int n_images, long long pixels_per_image, double* out,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ISA_CHECK_ARG(gt && pred && out && workspace, "isa_sbd: null pointer");
  ISA_CHECK_ARG(n_images > 0 && n_images <= 65535 && pixels_per_image > 0, "isa_sbd: n_images=%d pixels=%lld", n_images, pixels_per_image);
  const size_t need = isa_sbd_workspace_bytes(n_images);
  if (workspace_bytes < need) {
    isa_set_error("isa_sbd: workspace %zu < %zu", workspace_bytes, need);
    return ISA_ERR_WORKSPACE;
  }
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  ISA_CUDA(cudaMemsetAsync(workspace, 0, need, stream));
  long long gx = (pixels_per_image + 1023) / 1024;
  const long long cap = ((long long)di.num_sms * 8 + n_images - 1) / n_images;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  sbd_table_kernel<<<dim3((unsigned)gx, (unsigned)n_images), 256, 0, stream>>>(gt, pred, pixels_per_image, reinterpret_cast<unsigned*>(workspace));
  sbd_reduce_kernel<<<n_images, 256, 0, stream>>>(reinterpret_cast<const unsigned*>(workspace), out);
  ISA_CUDA(cudaGetLastError());
  return ISA_OK;
}
