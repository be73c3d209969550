// Shared helpers for the sm_100a kernels behind libisa_sm100.so.
// Everything in csrc/ is compiled by build.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ISA_OK 0
#define ISA_ERR_BAD_ARG (-1)
#define ISA_ERR_UNSUPPORTED (-2)
#define ISA_ERR_WORKSPACE (-3)

// Error plumbing: every entry point returns 0, a negative ISA_ERR_* for a
// rejected argument, or a positive cudaError_t.  The message is kept per
// thread and read back through isa_last_error().
void isa_set_error(const char* fmt, ...);

#define ISA_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      isa_set_error(__VA_ARGS__);           \
      return ISA_ERR_BAD_ARG;               \
    }                                       \
  } while (0)

#define ISA_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      isa_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                    cudaGetErrorString(_e));                                  \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

struct IsaDeviceInfo {
  int device;
  int num_sms;
  int max_smem_optin;
};
// Cached per device (cudaGetDeviceProperties is slow).
int isa_device_info(IsaDeviceInfo* out);

static inline size_t isa_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Group barrier between `expected` co-resident CTAs (cooperative launch
// guarantees residency).  `counter` must start at 0 and is monotonically
// increasing; `target` = expected * (number of barriers passed so far + 1).
__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    while (ld_acquire_u32(counter) < target) { __nanosleep(32); }
    __threadfence();  // acquire side: drop stale L1 lines before the CTA reads other CTAs' data
  }
  __syncthreads();
}
// Variant for barriers behind which a few CTAs work for a long time while the rest wait (the k-means++ candidate search:
// 35 of 256 CTAs): the waiters poll with RELAXED loads and back off, and fence once at the end.  ld.acquire.gpu compiles
// to LDG.STRONG + CCTL.IVALL; a waiter that invalidates its SM's L1 every ~100 ns slowed the working CTA on the same SM.
__device__ __forceinline__ void group_barrier_patient(unsigned* counter, unsigned target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    unsigned ns = 64;
    while (true) {
      unsigned v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;
    }
    __threadfence();
  }
  __syncthreads();
}
// Leaner variant (the cooperative-groups pattern): bar.sync orders the CTA's writes before thread 0, whose
// gpu-scope fence + release-arrive publishes them (fence cumulativity); no per-thread fence, no sleep in the poll.
__device__ __forceinline__ void group_barrier_lean(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (ld_acquire_u32(counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}
#endif
