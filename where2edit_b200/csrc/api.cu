// Library-wide plumbing of libw2e: error reporting, version, device query.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace w2e {

char* last_error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static int cached[64] = {};
  const int dev = current_device();
  if (dev < 0) return 148;  // B200; do not cache a failed query
  if (dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev < 64) cached[dev] = n;
  return n;
}

}  // namespace w2e

extern "C" int w2e_version(void) { return 100; }

extern "C" const char* w2e_last_error_string(void) { return w2e::last_error_buffer(); }

extern "C" int w2e_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  W2E_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  W2E_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count_out) *sm_count_out = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return W2E_OK;
}
