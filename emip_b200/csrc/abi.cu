// C-ABI plumbing shared by all entry points: error string, version, device probe.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include <stdarg.h>
#include <mutex>
#include <set>
#include <utility>

static thread_local char g_err[512] = "";

void emip_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- the library's only process-wide mutable state (besides diagnostic switches): two small caches behind one mutex ----
namespace {
std::mutex g_mu;
int g_sms[64];                                             // SM count per device ordinal (0 = not queried yet)
std::set<std::pair<const void*, int>> g_attr_done;         // (kernel, device) pairs whose shared-memory limit is raised
}  // namespace

int emip_num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sms[dev] = n > 0 ? n : 148;
  }
  return g_sms[dev];
}

int emip_func_max_smem(const void* func, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_attr_done.count({func, dev})) return EMIP_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    emip_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize, %d): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  g_attr_done.insert({func, dev});
  return EMIP_OK;
}

extern "C" const char* emip_last_error(void) { return g_err; }

extern "C" int emip_abi_version(void) { return EMIP_ABI_VERSION; }

extern "C" int emip_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    emip_set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    emip_set_error("emip_b200 is built for sm_100a only; device is sm_%d%d", major, minor);
    return EMIP_ENOSYS;
  }
  return EMIP_OK;
}

// Launch policy (include/emip_b200.h): programmatic dependent launch of the persistent tensor-core kernels
// (common.cuh::emip_launch_pdl).  Off unless the caller, who knows that ONE stream owns the GPU, switches it on.
static int g_pdl = 0;
int emip_pdl_enabled() { return g_pdl; }
extern "C" int emip_set_programmatic_launch(int on) {
  const int prev = g_pdl;
  g_pdl = on ? 1 : 0;
  return prev;
}
