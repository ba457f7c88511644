// Error plumbing and device queries of the C ABI (include/capr_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace capr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return CAPR_ERR_CUDA;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

int device_of(const void* p) {
  if (!p) return -1;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? attr.device : -1;
}

void DeviceGuard::enter(int device) {
  if (device < 0) return;
  if (cudaGetDevice(&prev) != cudaSuccess) {
    cudaGetLastError();
    prev = -1;
    return;
  }
  if (device != prev && cudaSetDevice(device) == cudaSuccess) switched = true;
}
DeviceGuard::DeviceGuard(const void* device_ptr) { enter(device_of(device_ptr)); }
DeviceGuard::DeviceGuard(int device) { enter(device); }
DeviceGuard::~DeviceGuard() {
  if (switched && prev >= 0) cudaSetDevice(prev);
}

}  // namespace capr

extern "C" {

int capr_abi_version(void) { return CAPR_ABI_VERSION; }
const char* capr_last_error(void) { return capr::g_err; }
int capr_device_sm_count(void) { return capr::sm_count(); }

int capr_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

}  // extern "C"
