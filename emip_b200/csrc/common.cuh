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

// Programmatic dependent launch (PDL) for the persistent tensor-core kernels: a kernel launched through emip_launch_pdl may start
// as soon as every CTA of the kernel in front of it on the stream has executed pdl_trigger() (or exited) and SMs free up; its
// prologue (barrier init, TMEM allocation, first descriptor fetches) then overlaps the tail of its predecessor, and pdl_wait()
// -- which every thread must pass before it touches memory the predecessor wrote -- returns once that grid has completed and
// flushed.  Both instructions are no-ops for a kernel launched the ordinary way, so kernels can adopt this one by one.  Captured
// into a CUDA graph the launch becomes a programmatic-dependency edge.  Used by gemm_tc_kernel, attn_fwd_tc_kernel and
// mlp_fused_kernel (one CTA per SM: nothing of the next kernel can start on an SM before this kernel's CTA there is gone).  Same-box
// A / B of the chain step (tools/pdl_ab.py, profiles/r5h_pdl_ab.txt): -0.8 % at 64 pairs, -1.3 % at 8.  Extending it to the small
// many-block kernels around them (operand splits, the prompt fusion's glue) was +1.4 % at 64 pairs (r5i: their blocks co-reside with
// the running persistent CTAs and wait there): not kept.  OFF by default (emip_set_programmatic_launch): with several streams on
// one GPU an early-launched CTA holds an SM another stream's kernel could use -- c4 with four clip streams lost 28 % (r5q).
int emip_pdl_enabled();     // abi.cu: the process-wide launch policy set through emip_set_programmatic_launch (default 0)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t emip_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = emip_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

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
