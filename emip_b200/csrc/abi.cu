// C-ABI plumbing shared by all entry points: error string, version, device probe.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void emip_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
