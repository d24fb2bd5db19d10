// Error handling, version and device queries of libynet_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ynet {

static thread_local char g_err[512] = "no error";
static int g_sm_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return YNET_E_CUDA;
}

int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      g_sm_count = 148;
  }
  return g_sm_count;
}

}  // namespace ynet

extern "C" {

int ynet_version(void) { return 100; }

const char* ynet_last_error_string(void) { return ynet::g_err; }

int ynet_device_info(int32_t* sms, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ynet::cuda_fail(e, "ynet_device_info");
  int n = 0, ma = 0, mi = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
  if (sms) *sms = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  if (ma != 10) {
    ynet::set_error("ynet_device_info: device is sm_%d%d, this library is built for sm_100a only", ma, mi);
    return YNET_E_ARCH;
  }
  return YNET_OK;
}

}  // extern "C"
