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
// FP32 pipe probe: 16 independent accumulator chains per thread, `iters` x 16 fused multiply-adds, scalar (FFMA) or
// packed (fma.rn.f32x2 -> FFMA2: two FMAs per instruction)
template <int PACKED>
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, float seed, float* sink) {
  float a = seed + threadIdx.x * 1e-3f, b = 1.0f - 1e-6f * blockIdx.x;
  if (PACKED) {
    unsigned long long acc[16], av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (unsigned long long)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t ^= acc[i];
    if (t == 0x1234567ull) *sink = 1.f;
  } else {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += acc[i];
    if (t == 12345.678f) *sink = t;
  }
}
}  // namespace

extern "C" {

// Diagnostics: FP32 multiply-adds per clock per SM sustained by scalar FFMA (packed = 0) or FFMA2 (packed = 1) with
// `warps_per_sm` resident warps; synchronises the device.  The register-resident GRU scan and the k-means E-step are
// sized against this number.  scratch = 8 bytes of device memory; sm_clock_mhz = the clock to convert time to cycles.
int isa_selftest_fma_rate(int packed, int warps_per_sm, float sm_clock_mhz, void* scratch, float* h_fma_per_clk_per_sm) {
  ISA_CHECK_ARG(warps_per_sm > 0 && warps_per_sm <= 64 && sm_clock_mhz > 0 && scratch && h_fma_per_clk_per_sm, "selftest_fma_rate: bad argument");
  IsaDeviceInfo di;
  int rc = isa_device_info(&di);
  if (rc) return rc;
  const int iters = 1 << 15;
  const int threads = 256;
  const int ctas_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
  const int thr = warps_per_sm * 32 < threads ? warps_per_sm * 32 : threads;
  cudaEvent_t e0, e1;
  ISA_CUDA(cudaEventCreate(&e0));
  ISA_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    ISA_CUDA(cudaEventRecord(e0, 0));
    if (packed) fma_probe_kernel<1><<<di.num_sms * ctas_per_sm, thr>>>(iters, 0.5f, reinterpret_cast<float*>(scratch));
    else fma_probe_kernel<0><<<di.num_sms * ctas_per_sm, thr>>>(iters, 0.5f, reinterpret_cast<float*>(scratch));
    ISA_CUDA(cudaEventRecord(e1, 0));
    ISA_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ISA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double fmas_per_sm = (double)iters * 16 * (packed ? 2 : 1) * (double)ctas_per_sm * thr;
  *h_fma_per_clk_per_sm = (float)(fmas_per_sm / ((double)best * 1e-3 * sm_clock_mhz * 1e6));
  return ISA_OK;
}

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
