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

extern "C" {

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
