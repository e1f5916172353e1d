// Shared helpers for the emip_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define EMIP_OK 0
#define EMIP_EINVAL (-22)
#define EMIP_ENOSYS (-38)
#define EMIP_ENOMEM (-12)

// Records a message retrievable through emip_last_error() (thread-local).
void emip_set_error(const char* fmt, ...);

#define EMIP_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      emip_set_error(__VA_ARGS__);               \
      return EMIP_EINVAL;                        \
    }                                            \
  } while (0)

// Launch-error check: never synchronises; returns the cudaError_t (positive) on failure.
#define EMIP_CHECK_LAUNCH(what)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      emip_set_error("%s: %s", what, cudaGetErrorString(e__));               \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

#define EMIP_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      emip_set_error("%s: %s", #call, cudaGetErrorString(e__));              \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

// SM count of the CURRENT device (cached per device; one process may drive several GPUs).
int emip_num_sms();
// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes) once per (kernel, device): the attribute is per device.
// Mutex-guarded; returns EMIP_OK or the cudaError_t.
int emip_func_max_smem(const void* func, int bytes);

static inline size_t emip_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
