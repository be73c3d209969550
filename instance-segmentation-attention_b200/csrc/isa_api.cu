// Library-level plumbing of libisa_sm100.so: error string, device info cache,
// version.  The boundary idiom mirrors the reference's one native op
// (/root/reference/code/lib/archs/modules/sru/cuda_functional.py:441-547):
// flat extern "C" entry points taking raw device pointers + sizes, launched on
// the caller's stream, all memory owned by the caller.
#include "isa_common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void isa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int isa_device_info(IsaDeviceInfo* out) {
  static thread_local IsaDeviceInfo cache[16];
  static thread_local bool valid[16] = {false};
  int dev = 0;
  ISA_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 16 && valid[dev]) { *out = cache[dev]; return ISA_OK; }
  IsaDeviceInfo di;
  di.device = dev;
  ISA_CUDA(cudaDeviceGetAttribute(&di.num_sms, cudaDevAttrMultiProcessorCount, dev));
  ISA_CUDA(cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int major = 0, minor = 0;
  ISA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ISA_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    isa_set_error("libisa_sm100 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return ISA_ERR_UNSUPPORTED;
  }
  if (dev >= 0 && dev < 16) { cache[dev] = di; valid[dev] = true; }
  *out = di;
  return ISA_OK;
}

namespace {
// `iters` grid barriers back to back; variant 0 = group_barrier (fence per thread, sleeping poll), 1 = group_barrier_lean
__global__ void barrier_probe_kernel(unsigned* counter, int iters, int variant, float* sink) {
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    if (variant == 0) group_barrier(counter, (unsigned)(i + 1) * gridDim.x);
    else group_barrier_lean(counter, (unsigned)(i + 1) * gridDim.x);
    acc += 1.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sink = acc;
}
}  // namespace

extern "C" {

// Diagnostics: mean device time (us) of one grid-wide barrier between ctas_per_sm * num_sms co-resident CTAs of
// `threads` threads -- the fixed cost every iteration of the persistent cooperative kernels (k-means, loss) pays.
// Synchronises the device; `scratch` = 8 bytes of device memory.
int isa_selftest_grid_barrier(int ctas_per_sm, int threads, int iters, int variant, void* scratch, float* h_us_per_barrier) {
  ISA_CHECK_ARG(ctas_per_sm > 0 && threads > 0 && threads <= 1024 && iters > 0 && scratch && h_us_per_barrier, "selftest_grid_barrier: bad argument");
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  unsigned* counter = reinterpret_cast<unsigned*>(scratch);
  float* sink = reinterpret_cast<float*>(counter + 1);
  cudaEvent_t e0, e1;
  ISA_CUDA(cudaEventCreate(&e0));
  ISA_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    ISA_CUDA(cudaMemset(counter, 0, 8));
    void* args[] = {(void*)&counter, (void*)&iters, (void*)&variant, (void*)&sink};
    ISA_CUDA(cudaEventRecord(e0, 0));
    ISA_CUDA(cudaLaunchCooperativeKernel((const void*)barrier_probe_kernel, dim3(di.num_sms * ctas_per_sm), dim3(threads), args, 0, 0));
    ISA_CUDA(cudaEventRecord(e1, 0));
    ISA_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ISA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *h_us_per_barrier = best * 1e3f / (float)iters;
  return ISA_OK;
}

const char* isa_last_error(void) { return g_err; }

int isa_version(void) { return 100; }

// Number of SMs of the current device (also the "is there a usable B200" probe).
int isa_num_sms(int* out) {
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  if (out) *out = di.num_sms;
  return ISA_OK;
}

}  // extern "C"
