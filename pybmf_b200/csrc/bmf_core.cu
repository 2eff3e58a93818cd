// Error plumbing and device queries for libbmf_b200.so.
#include <stdarg.h>
#include <string.h>

#include "bmf_common.cuh"

namespace bmf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_arg(const char* what) {
  set_error("%s", what);
  return BMF_E_ARG;
}

int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  set_error("%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int num_sms() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  cached = sms;
  return cached;
}

}  // namespace bmf

extern "C" int bmf_abi_version(void) { return BMF_ABI_VERSION; }

extern "C" const char* bmf_last_error(void) { return bmf::g_err; }

extern "C" int bmf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    bmf::set_error("bmf_device_info: no CUDA device (%s)", cudaGetErrorString(e));
    return BMF_E_NOGPU;
  }
  int sms = 0, maj = 0, min = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    bmf::set_error("bmf_device_info: device is sm_%d%d, this library is built for sm_100a only", maj, min);
    return BMF_E_NOGPU;
  }
  return 0;
}
